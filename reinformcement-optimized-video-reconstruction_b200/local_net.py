"""Drop-in for the reference `local_net` module: `LocalNetworkUNetNorm(freeze=False)`.

Same class name, constructor argument, `forward(x, context)` signature, parameter order and
72-key `state_dict` as rovr/local_net.py:7-72 (including the ten BatchNorm2d that the reference
constructs but never calls — rovr/local_net.py:13-37 vs :52-68 — so checkpoints load unchanged),
but the arithmetic runs in hand-written sm_100a kernels through librovr_b200.so:

    pack(x, context) -> NHWC bf16, 16 ch          (replaces cat + rearrange, :48-49)
    conv1..conv4 (+ReLU), 2x2 max-pool             tcgen05 implicit GEMM, pool kernel
    upconv1..3 (+ReLU)                             GEMM + pixel-shuffle store into the concat slice
    conv5..conv7 (+ReLU) on [up, skip] buffers     the torch.cat of :59,63,67 never materialises
    conv8 1x1 + sigmoid (+ optional fused L2)      fused tail kernel

and the backward pass (dgrad / wgrad on tensor cores, bias grads, pool scatter fused with the
skip-gradient add and ReLU mask) is one autograd.Function. Inputs / outputs at the module
boundary stay NCHW fp32 like the reference; internals are NHWC bf16 with fp32 accumulation and
fp32 master weights. There is no fallback path: a CPU tensor or a non-sm_100 device raises.
"""
import os

import torch
import torch.nn as nn

import ops

# (name, kind, cin, cout) in the reference's registration order (rovr/local_net.py:12-39)
_LAYERS = [
    ("conv1", "conv", 9, 64), ("bn1", "bn", 64, 64),
    ("conv2", "conv", 64, 128), ("bn2", "bn", 128, 128),
    ("conv3", "conv", 128, 256), ("bn3", "bn", 256, 256),
    ("conv4", "conv", 256, 512), ("bn4", "bn", 512, 512),
    ("upconv1", "up", 512, 256), ("bn_up1", "bn", 256, 256),
    ("conv5", "conv", 512, 256), ("bn5", "bn", 256, 256),
    ("upconv2", "up", 256, 128), ("bn_up2", "bn", 128, 128),
    ("conv6", "conv", 256, 128), ("bn6", "bn", 128, 128),
    ("upconv3", "up", 128, 64), ("bn_up3", "bn", 64, 64),
    ("conv7", "conv", 128, 64), ("bn7", "bn", 64, 64),
]

# parameters that take part in forward, in backward-completion order (decoder first): the order
# in which their gradients become final, used for bucketed all-reduce overlap.
_GRAD_ORDER = ["conv8", "conv7", "upconv3", "conv6", "upconv2", "conv5", "upconv1",
               "conv4", "conv3", "conv2", "conv1"]
_DECODER = _GRAD_ORDER[:7]
_ENCODER = _GRAD_ORDER[7:]
# all-reduce buckets, in the order their gradients become final: the decoder (8.95 MB, ready after upconv1's
# weight gradient), conv4 alone (4.72 MB, ready one kernel later) and the thin rest of the encoder (1.50 MB): only
# the last, smallest bucket's all-reduce is exposed at the end of the step (round 1 had two buckets and the whole
# 6.2 MB encoder one at the end)
_BUCKETS = [_DECODER, ["conv4"], ["conv3", "conv2", "conv1"]]
if os.environ.get("ROVR_DP_BUCKETS") == "2":      # A/B switch: the round-1 layout (decoder, whole encoder)
    _BUCKETS = [_DECODER, _ENCODER]


class _PackedWeights:
    """bf16 GEMM-operand copies of the fp32 master weights, refreshed when a parameter changes."""

    def __init__(self):
        self._cache = {}

    def get(self, name, param, kind, for_dgrad):
        key = (name, for_dgrad)
        ver = (param.data_ptr(), param._version)
        hit = self._cache.get(key)
        if hit is not None and hit[0] == ver:
            return hit[1]
        w = param.detach()
        wk = ops.repack_convT2x2(w, for_dgrad) if kind == "up" else ops.repack_conv3x3(w, for_dgrad)
        self._cache[key] = (ver, wk)
        return wk

    def refresh(self, layers, with_dgrad):
        """layers: [(name, param, kind)]. Re-packs every stale operand copy (forward operands, and the
        data-gradient operands of all layers but the first when `with_dgrad`) with ONE launch, so that
        the per-layer get() calls of the step only hit the cache."""
        todo = []
        for i, (name, param, kind) in enumerate(layers):
            ver = (param.data_ptr(), param._version)
            for for_dgrad in ((False, True) if (with_dgrad and i > 0) else (False,)):
                hit = self._cache.get((name, for_dgrad))
                if hit is None or hit[0] != ver:
                    todo.append((name, for_dgrad, ver, param.detach(), kind))
        if not todo:
            return
        outs = ops.repack_batch([(w, kind, fd) for _, fd, _, w, kind in todo])
        for (name, fd, ver, _, _), wk in zip(todo, outs):
            self._cache[(name, fd)] = (ver, wk)


def _flat_bucket(params, names, device):
    """One flat fp32 buffer holding weight+bias grads of `names`; returns (flat, {pname: view})."""
    total = sum(params[n + ".weight"].numel() + params[n + ".bias"].numel() for n in names)
    flat = torch.empty(total, dtype=torch.float32, device=device)
    views, off = {}, 0
    for n in names:
        for suffix in (".weight", ".bias"):
            p = params[n + suffix]
            views[n + suffix] = flat[off:off + p.numel()].view(p.shape)
            off += p.numel()
    return flat, views


def _forward_impl(net, x, context, target, P, with_dgrad):
    """The forward launch sequence. P: {live parameter name: tensor}. Returns (out, loss | None, acts)."""
    dev = x.device
    B, _, H, W = x.shape
    bf = torch.bfloat16
    pk = net._packed
    # all stale bf16 operand copies (forward and, if a backward will follow, data-gradient) in one launch
    pk.refresh([(n, P[n + ".weight"], k) for n, k, _, _ in _LAYERS if k != "bn"], with_dgrad)

    def wk(name, kind="conv"):
        return pk.get(name, P[name + ".weight"], kind, False)

    a = {}
    a["in16"] = ops.pack_nchw([x, context.reshape(B, 6, H, W)], 16)
    a["cat7"] = torch.empty((B, H, W, 128), dtype=bf, device=dev)
    a["cat6"] = torch.empty((B, H // 2, W // 2, 256), dtype=bf, device=dev)
    a["cat5"] = torch.empty((B, H // 4, W // 4, 512), dtype=bf, device=dev)
    x1, x2, x3 = a["cat7"][..., 64:], a["cat6"][..., 128:], a["cat5"][..., 256:]
    def conv_pool(src, name, dst, pooled):
        # conv + ReLU + MaxPool2d(2,2) in one kernel when the halo tiling applies (W >= 8, H >= 16)
        if src.shape[1] >= 16 and src.shape[2] >= 8:
            ops.conv3x3_fprop(src, wk(name), P[name + ".bias"], dst, pooled=pooled)
        else:
            ops.conv3x3_fprop(src, wk(name), P[name + ".bias"], dst)
            ops.maxpool_fwd(dst, pooled, 2)

    a["p1"] = torch.empty((B, H // 2, W // 2, 64), dtype=bf, device=dev)
    conv_pool(a["in16"], "conv1", x1, a["p1"])
    a["p2"] = torch.empty((B, H // 4, W // 4, 128), dtype=bf, device=dev)
    conv_pool(a["p1"], "conv2", x2, a["p2"])
    a["p3"] = torch.empty((B, H // 8, W // 8, 256), dtype=bf, device=dev)
    conv_pool(a["p2"], "conv3", x3, a["p3"])
    a["x4"] = torch.empty((B, H // 8, W // 8, 512), dtype=bf, device=dev)
    ops.conv3x3_fprop(a["p3"], wk("conv4"), P["conv4.bias"], a["x4"])

    ops.convT2x2_fprop(a["x4"], wk("upconv1", "up"), P["upconv1.bias"], a["cat5"][..., :256])
    a["y5"] = torch.empty((B, H // 4, W // 4, 256), dtype=bf, device=dev)
    ops.conv3x3_fprop(a["cat5"], wk("conv5"), P["conv5.bias"], a["y5"])
    ops.convT2x2_fprop(a["y5"], wk("upconv2", "up"), P["upconv2.bias"], a["cat6"][..., :128])
    a["y6"] = torch.empty((B, H // 2, W // 2, 128), dtype=bf, device=dev)
    ops.conv3x3_fprop(a["cat6"], wk("conv6"), P["conv6.bias"], a["y6"])
    ops.convT2x2_fprop(a["y6"], wk("upconv3", "up"), P["upconv3.bias"], a["cat7"][..., :64])
    a["y7"] = torch.empty((B, H, W, 64), dtype=bf, device=dev)
    # conv7 + ReLU + conv8 (1x1) + sigmoid (+ L2 loss) in one kernel: the conv7 epilogue thread owns
    # a whole pixel of y7, so the 64 -> 3 projection needs no second pass over it
    out, loss = ops.conv3x3_fprop_tail(a["cat7"], wk("conv7"), P["conv7.bias"], a["y7"], P["conv8.weight"],
                                       P["conv8.bias"], target)
    a["out"] = out
    return out, loss, a


_SIDE_STREAMS = {}


def _wgrad_side_stream(dev):
    """Weight-gradient launches (split-K kernel + its reduction) go to a forked stream when the step is launched
    eagerly: a layer's weight gradient and data gradient are independent, and with two hardware queues the tail of one
    persistent kernel is filled by the CTAs of the other and the small reductions overlap the next tensor-bound
    kernel. Measured on one B200 (alternating A/B, profiles/r02_wgrad_stream_ab.log): eager step 3.19 -> 3.08 ms, but a
    CUDA-graph replay of the same fork / join gets 1 % SLOWER (3.19 -> 3.23), so it is not used under stream capture.
    ROVR_WGRAD_STREAM=0 / 1 forces it off / on everywhere."""
    mode = os.environ.get("ROVR_WGRAD_STREAM")
    if mode == "0" or (mode != "1" and torch.cuda.is_current_stream_capturing()):
        return None
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=dev)
    return _SIDE_STREAMS[key]


def _backward_impl(net, a, P, target, g_out, g_loss, buckets=None):
    """The backward launch sequence. Writes every parameter gradient into three flat fp32 buckets
    (`_BUCKETS`; `buckets` = ([flat, ...], views) to reuse static ones) and returns {parameter name: gradient view}."""
    dev = a["out"].device
    bf = torch.bfloat16
    pk = net._packed
    if g_out is not None:
        g_out = g_out.contiguous()
    B, H, W, _ = a["y7"].shape

    def wd(name, kind="conv"):
        return pk.get(name, P[name + ".weight"], kind, True)

    if buckets is None:
        flats, G = _make_buckets(P, dev)
    else:
        flats, G = buckets
    net._last_buckets = tuple(flats)

    def el(t):
        return torch.empty_like(t)

    side = _wgrad_side_stream(dev)

    def wgrad(fn, *args):
        if side is None:
            return fn(*args)
        side.wait_stream(torch.cuda.current_stream(dev))      # its operands were produced on the main stream
        with torch.cuda.stream(side):
            fn(*args)

    def join():
        if side is not None:
            torch.cuda.current_stream(dev).wait_stream(side)

    # ---- tail: conv8 + sigmoid (+L2) ----
    g7 = el(a["y7"])
    use_loss = target is not None and g_loss is not None
    ops.tail_bwd(a["y7"], P["conv8.weight"], a["out"], g7, G["conv8.weight"], G["conv8.bias"],
                 gout=g_out, target=target if use_loss else None,
                 mse_scale=2.0 / a["out"].numel(),
                 gloss=g_loss.contiguous() if use_loss else None, db7=G["conv7.bias"])
    # ---- conv7 (its bias gradient came out of the tail kernel) ----
    wgrad(ops.conv3x3_wgrad, g7, a["cat7"], G["conv7.weight"])
    gcat7 = el(a["cat7"])
    # the ReLU mask is only needed on the up-conv half: the skip half (x1) is masked by pool1's backward
    ops.conv3x3_dgrad(g7, wd("conv7"), gcat7, mask=a["cat7"], mask_cols=64, colsum=G["upconv3.bias"])
    # ---- upconv3 ----
    gu3 = gcat7[..., :64]
    wgrad(ops.convT2x2_wgrad, gu3, a["y6"], G["upconv3.weight"])
    g6 = el(a["y6"])
    ops.convT2x2_dgrad(gu3, wd("upconv3", "up"), g6, mask=a["y6"], colsum=G["conv6.bias"])
    # ---- conv6 ----
    wgrad(ops.conv3x3_wgrad, g6, a["cat6"], G["conv6.weight"])
    gcat6 = el(a["cat6"])
    ops.conv3x3_dgrad(g6, wd("conv6"), gcat6, mask=a["cat6"], mask_cols=128, colsum=G["upconv2.bias"])
    # ---- upconv2 ----
    gu2 = gcat6[..., :128]
    wgrad(ops.convT2x2_wgrad, gu2, a["y5"], G["upconv2.weight"])
    g5 = el(a["y5"])
    ops.convT2x2_dgrad(gu2, wd("upconv2", "up"), g5, mask=a["y5"], colsum=G["conv5.bias"])
    # ---- conv5 ----
    wgrad(ops.conv3x3_wgrad, g5, a["cat5"], G["conv5.weight"])
    gcat5 = el(a["cat5"])
    ops.conv3x3_dgrad(g5, wd("conv5"), gcat5, mask=a["cat5"], mask_cols=256, colsum=G["upconv1.bias"])
    # ---- upconv1 ----
    gu1 = gcat5[..., :256]
    wgrad(ops.convT2x2_wgrad, gu1, a["x4"], G["upconv1.weight"])
    g4 = el(a["x4"])
    ops.convT2x2_dgrad(gu1, wd("upconv1", "up"), g4, mask=a["x4"], colsum=G["conv4.bias"])
    join()
    net._bucket_ready(0, flats[0])
    # ---- conv4 ----
    wgrad(ops.conv3x3_wgrad, g4, a["p3"], G["conv4.weight"])
    if len(flats) == 3:
        join()
        net._bucket_ready(1, flats[1])
    gp3 = el(a["p3"])
    ops.conv3x3_dgrad(g4, wd("conv4"), gp3)
    g3 = torch.empty((B, H // 4, W // 4, 256), dtype=bf, device=dev)
    ops.maxpool_bwd(a["cat5"][..., 256:], gp3, g3, 2, gskip=gcat5[..., 256:], relu_mask=True,
                    colsum=G["conv3.bias"])             # bias gradient from the same pass over g3
    # ---- conv3 ----
    wgrad(ops.conv3x3_wgrad, g3, a["p2"], G["conv3.weight"])
    gp2 = el(a["p2"])
    ops.conv3x3_dgrad(g3, wd("conv3"), gp2)
    g2 = torch.empty((B, H // 2, W // 2, 128), dtype=bf, device=dev)
    ops.maxpool_bwd(a["cat6"][..., 128:], gp2, g2, 2, gskip=gcat6[..., 128:], relu_mask=True,
                    colsum=G["conv2.bias"])
    # ---- conv2 ----
    wgrad(ops.conv3x3_wgrad, g2, a["p1"], G["conv2.weight"])
    gp1 = el(a["p1"])
    ops.conv3x3_dgrad(g2, wd("conv2"), gp1)
    g1 = torch.empty((B, H, W, 64), dtype=bf, device=dev)
    ops.maxpool_bwd(a["cat7"][..., 64:], gp1, g1, 2, gskip=gcat7[..., 64:], relu_mask=True,
                    colsum=G["conv1.bias"])
    # ---- conv1 (input needs no gradient: no dgrad, as in the reference's autograd graph) ----
    wgrad(ops.conv3x3_wgrad, g1, a["in16"], G["conv1.weight"])
    join()
    net._bucket_ready(len(flats) - 1, flats[-1])
    net._buckets_wait()
    return G


def _make_buckets(P, dev):
    flats, G = [], {}
    for names in _BUCKETS:
        flat, views = _flat_bucket(P, names, dev)
        flats.append(flat)
        G.update(views)
    return flats, G


class _LocalNetFunction(torch.autograd.Function):
    """forward: (x, context, target_or_None, *live_params) -> (y_hat, loss_or_zero)."""

    @staticmethod
    def forward(ctx, net, x, context, target, *plist):
        P = dict(zip(net._live_names, plist))
        out, loss, a = _forward_impl(net, x, context, target, P, any(ctx.needs_input_grad))
        ctx.net = net
        ctx.acts = a
        ctx.params = P
        ctx.target = target
        ctx.set_materialize_grads(False)
        if loss is None:
            loss = out.new_zeros(())
            ctx.mark_non_differentiable(loss)
        return out, loss

    @staticmethod
    def backward(ctx, g_out, g_loss):
        net, a, P, target = ctx.net, ctx.acts, ctx.params, ctx.target
        if g_out is None and (g_loss is None or target is None):
            return (None,) * (4 + len(net._live_names))
        G = _backward_impl(net, a, P, target, g_out, g_loss)
        ctx.acts = None
        grads = tuple(G[n] if P[n].requires_grad else None for n in net._live_names)
        return (None, None, None, None) + grads


class LocalNetworkUNetNorm(nn.Module):
    """U-Net inpainting network; see the module docstring. Reference: rovr/local_net.py:7-72."""

    def __init__(self, freeze=False):
        super().__init__()
        for name, kind, cin, cout in _LAYERS:
            if kind == "conv":
                layer = nn.Conv2d(cin, cout, kernel_size=3, padding=1)
            elif kind == "up":
                layer = nn.ConvTranspose2d(cin, cout, kernel_size=2, stride=2)
            else:
                layer = nn.BatchNorm2d(cout)  # registered for state_dict parity; never applied
            setattr(self, name, layer)
        self.conv8 = nn.Conv2d(64, 3, kernel_size=1)
        self.maxpool = nn.MaxPool2d(kernel_size=2, stride=2)  # attribute kept for API parity
        if freeze:
            for param in self.parameters():
                param.requires_grad = False
        self._live_names = [n + s for n in
                            ["conv1", "conv2", "conv3", "conv4", "upconv1", "conv5", "upconv2",
                             "conv6", "upconv3", "conv7", "conv8"] for s in (".weight", ".bias")]
        self._packed = _PackedWeights()
        self._grad_bucket_hook = None  # set by data_parallel.GradientBuckets
        self._grad_bucket_wait = None
        self._grad_bucket_reduce = None
        self._last_buckets = None

    # -- data-parallel hook points ---------------------------------------------------------------
    def _bucket_ready(self, index, flat):
        if self._grad_bucket_hook is not None:
            self._grad_bucket_hook(index, flat)

    def _buckets_wait(self):
        if self._grad_bucket_wait is not None:
            self._grad_bucket_wait()

    # -- execution -------------------------------------------------------------------------------
    def _live_params(self):
        sd = dict(self.named_parameters())
        return [sd[n] for n in self._live_names]

    def _run(self, x, context, target):
        if not x.is_cuda:
            raise RuntimeError("LocalNetworkUNetNorm (B200) needs CUDA tensors: there is no CPU path")
        B, C, H, W = x.shape
        if C != 3 or context.shape != (B, 2, 3, H, W):
            raise ValueError(f"expected x [b,3,h,w] and context [b,2,3,h,w], got {tuple(x.shape)} "
                             f"{tuple(context.shape)}")
        if H % 8 or W % 8:
            raise ValueError("H and W must be multiples of 8 (three 2x2 poolings)")
        x = x.float().contiguous()
        context = context.float().contiguous()
        if target is not None:
            target = target.float().contiguous()
        with torch.cuda.device(x.device):   # launches go to the tensors' GPU, not the thread's current one
            return _LocalNetFunction.apply(self, x, context, target, *self._live_params())

    def forward(self, x, context):
        out, _ = self._run(x, context, None)
        return out

    def forward_with_mse(self, x, context, target):
        """(y_hat, mean((y_hat - target)^2)) with the loss and its gradient fused into the tail
        kernel: the `mse_loss_fn(y_hat, target)` of rovr/train_local_net_unet.py:107."""
        return self._run(x, context, target)

    def forward_with_loss(self, x, context, target, gamma, lpips_fn, normalize=False):
        """The full loss of rovr/train_local_net_unet.py:105-113: returns (y_hat, mse, lpips, total) with
        total = gamma * mse + (1 - gamma) * lpips_fn(y_hat, target).mean(); `lpips_fn` is a
        lpips_vgg.LPIPS. `total.backward()` runs the LPIPS dgrad chain into y_hat and then this network's
        backward with both the perceptual gradient and the fused L2 gradient."""
        y_hat, mse = self._run(x, context, target)
        lp = lpips_fn(y_hat, target, normalize=normalize).mean()
        return y_hat, mse, lp, mse * gamma + lp * (1.0 - gamma)


class GraphedTrainingStep:
    """forward + fused L2 loss + backward of a LocalNetworkUNetNorm captured ONCE into a CUDA graph.

    A step is a fixed sequence of ~56 kernel launches of 20-400 us each; replaying it as a graph
    removes the per-launch CPU cost and the gaps between kernels. Usage
    (the loop of rovr/train_local_net_unet.py:102-116, `zero_grad()` included):

        step = GraphedTrainingStep(net, frame, context, target)     # captures with these shapes
        for frame, context, target in batches:
            optimizer.zero_grad()
            loss = step(frame, context, target)      # copies into the static inputs, replays
            optimizer.step()                         # net.<param>.grad are this step's gradients

    The launch sequence is captured WITHOUT the autograd engine (`_forward_impl` / `_backward_impl`
    called directly): every gradient is written into three static flat fp32 buckets and each replay
    re-binds `param.grad` to its view of them, so the step survives `zero_grad(set_to_none=True)`
    (which drops `.grad`) and the gradients provably alias the buckets that are all-reduced.

    The bf16 operand copies of the weights are re-packed inside the graph (`repack_weights=True`, the
    default: a training loop's optimizer changes the fp32 masters every step, one launch);
    with False they are packed once before capture — only valid while the weights never change.

    With data_parallel.GradientBuckets installed the three NCCL all-reduces are captured INSIDE the
    graph on a forked stream: the decoder and conv4 buckets are reduced while the rest of backward is
    still running, only the 1.5 MB tail bucket at the end (`allreduce_mode == "captured-overlapped"`). If the
    collective cannot be captured (backend without graph support) the buckets are all-reduced right
    after each replay instead (`"after-replay"`).
    """

    def __init__(self, net, x, context, target, repack_weights=True, warmup=2, capture_collectives=True,
                 lpips_fn=None, gamma=1.0, normalize=False, input_sets=1):
        """lpips_fn (a lpips_vgg.LPIPS) adds the perceptual term of rovr/train_local_net_unet.py:109-113 to
        the captured step: loss = gamma * mse + (1 - gamma) * lpips(y_hat, target).mean(); `gamma` may be
        changed per step with `set_gamma()` (the reference anneals it, :111).

        input_sets=2 captures the step TWICE, on two sets of static inputs / outputs (`input_slots[k]`,
        `replay(k)`), sharing the weights and the gradient buckets: a feeder can then land batch i+1 straight in
        the other set while batch i is being computed (`DeviceFeeder(..., slots=step.input_slots)`), so that no
        staging kernel sits between two replays on the compute stream, and the loss of replay k stays valid
        until the next replay of the same set (it can be read back on a side stream). Measured: 3 staging
        kernels + a same-stream 4-byte read-back between replays cost 0.25 ms of a 3.2 ms step."""
        self.net = net
        self.repack_weights = repack_weights
        self.lpips_fn = lpips_fn
        dev = x.device
        self.input_slots = [(x.clone(), context.clone(), target.clone()) for _ in range(max(1, input_sets))]
        self.x, self.context, self.target = self.input_slots[0]
        named = dict(net.named_parameters())
        self.P = {n: named[n] for n in net._live_names}
        Pd = {n: p.detach() for n, p in self.P.items()}
        flats, G = _make_buckets(Pd, dev)
        self.buckets = tuple(flats)
        self.grad_views = G
        self._g_loss = torch.ones((), dtype=torch.float32, device=dev)        # d total / d mse  (= gamma)
        self._g_lpips = torch.zeros(x.shape[0], dtype=torch.float32, device=dev)  # d total / d lpips[n] (= (1-gamma)/N)
        self.lpips = None                                                        # static per-image LPIPS values
        if lpips_fn is not None:
            self.set_gamma(gamma)
        hooks = (net._grad_bucket_hook, net._grad_bucket_wait, net._grad_bucket_reduce)
        self._reduce_after = None
        self.allreduce_mode = "none"

        self._lpips_vals = [None] * len(self.input_slots)

        def run(k=0):
            xi, ci, ti = self.input_slots[k]
            y, loss, acts = _forward_impl(net, xi, ci, ti, Pd, True)
            g_out = None
            if lpips_fn is not None:
                import lpips_vgg
                f32 = lpips_fn.precision == "fp32x"
                fwd = lpips_vgg._forward_impl_f32 if f32 else lpips_vgg._forward_impl
                bwd = lpips_vgg._backward_impl_f32 if f32 else lpips_vgg._backward_impl
                self._lpips_vals[k], saved = fwd(lpips_fn, y, ti, normalize, True)
                self.lpips = self._lpips_vals[k]
                g_out = bwd(lpips_fn, saved, self._g_lpips)
            _backward_impl(net, acts, Pd, ti, g_out, self._g_loss, buckets=(flats, G))
            return y, loss

        import _native
        with torch.cuda.device(dev):
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(max(1, warmup)):     # packs the weights, warms the allocator and NCCL
                    run()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            want_collectives = hooks[0] is not None and capture_collectives
            for attempt in ((True, False) if want_collectives else (False,)):
                if not attempt:
                    net._grad_bucket_hook = net._grad_bucket_wait = None     # capture compute only
                if repack_weights:
                    net._packed._cache.clear()
                n0 = _native.lib.rovr_launch_count()
                self.graph = torch.cuda.CUDAGraph()
                try:
                    with torch.cuda.graph(self.graph, stream=side):     # the warm-up's stream: its scratch workspace is reused
                        self.y, self.loss = run()
                except Exception as exc:      # noqa: BLE001 — a collective that cannot be captured
                    if not attempt:
                        net._grad_bucket_hook, net._grad_bucket_wait = hooks[0], hooks[1]
                        raise
                    self._capture_error = repr(exc)
                    torch.cuda.synchronize(dev)
                    continue
                self.launches_per_step = int(_native.lib.rovr_launch_count() - n0)
                if hooks[0] is not None:
                    self.allreduce_mode = "captured-overlapped" if attempt else "after-replay"
                    self._reduce_after = None if attempt else hooks[2]
                break
            # further input sets: the same step captured again on its own static inputs / outputs
            self.graphs, self.ys, self.losses = [self.graph], [self.y], [self.loss]
            for k in range(1, len(self.input_slots)):
                if repack_weights:
                    net._packed._cache.clear()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side):
                    yk, lk = run(k)
                self.graphs.append(g)
                self.ys.append(yk)
                self.losses.append(lk)
        net._grad_bucket_hook, net._grad_bucket_wait = hooks[0], hooks[1]
        self._bind_grads()

    def set_gamma(self, gamma):
        """Loss weights of the next replays: total = gamma * mse + (1 - gamma) * mean(lpips) (two tiny fills,
        stream-ordered before the replay)."""
        self._g_loss.fill_(float(gamma))
        self._g_lpips.fill_((1.0 - float(gamma)) / self._g_lpips.numel())

    def close(self):
        """Destroy the captured graph. With NCCL collectives captured inside it this MUST happen before
        `dist.destroy_process_group()`: NCCL does not tear a communicator down while a CUDA graph that
        captured its collectives is alive (the teardown blocks forever)."""
        if getattr(self, "graph", None) is not None:
            torch.cuda.synchronize()
            for g in getattr(self, "graphs", [self.graph]):
                g.reset()
            self.graph = None
            self.graphs = []

    def _bind_grads(self):
        for n, p in self.P.items():
            if p.requires_grad:
                p.grad = self.grad_views[n]

    def grads_alias_buckets(self):
        """True iff every live `param.grad` is a view into the static buckets (checked by the tests)."""
        spans = [(b.data_ptr(), b.data_ptr() + b.numel() * 4) for b in self.buckets]
        return all(p.grad is not None and any(lo <= p.grad.data_ptr() < hi for lo, hi in spans)
                   for p in self.P.values() if p.requires_grad)

    def _load(self, dst, src):
        if src is None:
            return
        if src.dtype == torch.uint8:      # frames as decoded (feeder.DeviceFeeder(to_float=False)): ToTensor straight
            ops.u8_to_f32(src.contiguous(), out=dst)      # into the graph's static input, no fp32 staging copy
        else:
            dst.copy_(src, non_blocking=True)

    def replay(self, k=0):
        """Replay the step on input set k (whose tensors `input_slots[k]` the caller / feeder has filled, ordered
        before this call on the current stream). Returns that set's static loss tensor."""
        if self.graph is None:
            raise RuntimeError("GraphedTrainingStep.close() has been called")
        self.graphs[k].replay()
        if self._reduce_after is not None:
            self._reduce_after(self.buckets)
        self._bind_grads()
        self.y, self.loss, self.lpips = self.ys[k], self.losses[k], self._lpips_vals[k]
        return self.loss

    def __call__(self, x=None, context=None, target=None):
        if self.graph is None:
            raise RuntimeError("GraphedTrainingStep.close() has been called")
        self._load(self.x, x)
        self._load(self.context, context)
        self._load(self.target, target)
        self.graph.replay()
        if self._reduce_after is not None:
            self._reduce_after(self.buckets)                      # NCCL all-reduce (AVG) of the buckets
        self._bind_grads()      # the reference loop's zero_grad() sets .grad to None: re-attach the static views
        return self.loss
