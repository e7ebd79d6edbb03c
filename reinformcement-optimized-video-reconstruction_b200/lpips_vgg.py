"""B200 replacement for `lpips.LPIPS(net='vgg')`, the perceptual loss of the reference's training
steps (rovr/train_local_net_unet.py:91,109-113: `lpips_loss_fn(y_hat, target).mean()`, mixed with
the MSE by gamma; rovr/rovr.py:54,84,255: `self.lpips(y_hat, org_images, normalize=True)`, also the
RL reward). SURVEY §8f-1.

    from lpips_vgg import LPIPS
    lpips_loss_fn = LPIPS(net='vgg').to(device)
    lpips_loss = lpips_loss_fn(y_hat, target).mean()

The `lpips` package is not in this image and its trained weights need a download, so this is a
RESTATEMENT of its published v0.1 algorithm (PARITY UNPINNED against the package; pinned against
torchvision's VGG16 structure through oracle.lpips_vgg and tests/golden/lpips.npz) with random
weights by default; `load_state_dict` accepts the package's key layout (net.slice{k}.{idx}.*,
lin{k}.model.1.weight, scaling_layer.shift/scale).

Execution: both images are packed into ONE 2N batch (ScalingLayer fused into the pack kernel); the
13 VGG16 convolutions run on the tcgen05 implicit-GEMM engine (bf16 NHWC, bias + ReLU in the
epilogue, the 2x2 max-pool after relu1_2 .. relu4_3 emitted by the same epilogue); each tap's
distance (unit-normalise, weighted squared difference, spatial mean) AND its gradient w.r.t. the
first image's features come out of one pass over the features. Backward (to in0 only — the target
carries no gradient and VGG is frozen, so there is no weight gradient) is the dgrad chain of the
same engine; a tap's gradient joins the chain as the "skip" input of the fused pool-backward
kernel, exactly like a U-Net skip connection.
"""
import torch
import torch.nn as nn

import ops
from _blocks import BF, PackedWeights

# torchvision vgg16().features indices of the convolutions of each LPIPS slice (lpips/pretrained_networks.py)
_SLICES = [[(0, 3, 64), (2, 64, 64)], [(5, 64, 128), (7, 128, 128)],
           [(10, 128, 256), (12, 256, 256), (14, 256, 256)],
           [(17, 256, 512), (19, 512, 512), (21, 512, 512)],
           [(24, 512, 512), (26, 512, 512), (28, 512, 512)]]
_CHNS = [64, 128, 256, 512, 512]


def _forward_impl(mod, in0, in1, normalize, want_grad):
    """Returns (val [N] fp32, saved) — saved holds what the backward chain needs (None if not want_grad)."""
    N, _, H, W = in0.shape
    dev = in0.device
    pk = mod._packed
    shift = mod._shift_host
    scale = mod._scale_host
    x = ops.lpips_pack(in0, in1, shift, scale, normalize)
    inputs, feats = [], []
    for k, convs in enumerate(_SLICES):
        for j, (idx, cin, cout) in enumerate(convs):
            conv = getattr(getattr(mod.net, f"slice{k + 1}"), str(idx))
            wk = pk.get((k, idx, "f"), conv.weight, lambda t: ops.repack_conv3x3(t, False))
            B2, h, w, _ = x.shape
            y = torch.empty((B2, h, w, cout), dtype=BF, device=dev)
            last = j == len(convs) - 1
            inputs.append(x)
            if last and k < 4:
                pooled = torch.empty((B2, h // 2, w // 2, cout), dtype=BF, device=dev)
                if h >= 16 and w >= 8:                    # halo tiling: ReLU + MaxPool2d(2, 2) from the conv epilogue
                    ops.conv3x3_fprop(x, wk, conv.bias, y, pooled=pooled)
                else:
                    ops.conv3x3_fprop(x, wk, conv.bias, y)
                    ops.maxpool_fwd(y, pooled, 2)
                feats.append(y)
                x = pooled
            else:
                ops.conv3x3_fprop(x, wk, conv.bias, y)
                x = y
                if last:
                    feats.append(y)
    partials, grads = [], []
    for k, f in enumerate(feats):
        lin_w = getattr(mod, f"lin{k}").model[1].weight
        p, g = ops.lpips_head(f, lin_w.detach().reshape(-1), want_grad)
        partials.append(p)
        grads.append(g)
    val = ops.lpips_finalize(partials, [f.shape[1] * f.shape[2] for f in feats], N)
    saved = (inputs, feats, grads, N, normalize) if want_grad else None
    return val, saved


def _backward_impl(mod, saved, gval):
    """gval: [N] fp32 upstream gradient per image -> d/d in0 [N, 3, H, W] fp32."""
    inputs, feats, grads, N, normalize = saved
    pk = mod._packed
    dev = gval.device
    g = grads[4]                                          # d/d relu5_3 (already masked by the feature's ReLU)
    ci = len(inputs)
    for k in range(4, -1, -1):
        convs = _SLICES[k]
        for j in range(len(convs) - 1, -1, -1):
            idx, cin, cout = convs[j]
            ci -= 1
            conv = getattr(getattr(mod.net, f"slice{k + 1}"), str(idx))
            wd = pk.get((k, idx, "d"), conv.weight, lambda t: ops.repack_conv3x3(t, True))
            xin = inputs[ci][:N]
            dx = torch.empty(xin.shape, dtype=BF, device=dev)
            # ReLU backward of the producing convolution = mask by its (post-ReLU) output, which is this
            # convolution's input; the first convolution of a slice reads a pooled tensor, whose ReLU mask
            # the pool-backward kernel applies on the un-pooled activation instead
            ops.conv3x3_dgrad(g, wd, dx, mask=xin if j > 0 else None)
            g = dx
        if k > 0:
            f = feats[k - 1][:N]
            g = ops.maxpool_bwd(f, g, torch.empty(f.shape, dtype=BF, device=dev), 2, gskip=grads[k - 1], relu_mask=True)
    return ops.lpips_unpack_grad(g, gval.contiguous(), mod._shift_host, mod._scale_host, normalize)


# ---- precision="fp32x": the same network on the emulated-fp32 tensor-core path (csrc/fp32x.cuh) ---------------
# fp32 NHWC activations; every convolution is the 6-term split-bf16 product (forward) / 3-term (data gradient)
# stacked along the channel dimension, fp32 epilogue. Removes the ReLU-sign noise of bf16 operands from the pixel
# gradient d lpips / d in0 at ~6x the convolution cost.
def _pad16(c):
    return (c + 15) // 16 * 16


def _forward_impl_f32(mod, in0, in1, normalize, want_grad):
    N, _, H, W = in0.shape
    dev = in0.device
    pk = mod._packed
    x = ops.lpips_pack_f32(in0, in1, mod._shift_host, mod._scale_host, normalize)        # [2N, H, W, 4] fp32
    acts, feats = [], []
    pool = None
    for k, convs in enumerate(_SLICES):
        for j, (idx, cin, cout) in enumerate(convs):
            conv = getattr(getattr(mod.net, f"slice{k + 1}"), str(idx))
            cbi = ops.pad8(cin)
            wk = pk.get((k, idx, "f6"), conv.weight, lambda t: ops.repack_conv3x3(ops.split_weights(t, 1, 6, cbi), False))
            xs = ops.split_stack(x, 6, cbi, pool=pool)                   # MaxPool2d(2, 2) of the previous slice fused in
            pool = None
            B2, h, w, _ = xs.shape
            y = torch.empty((B2, h, w, cout), dtype=torch.float32, device=dev)
            ops.conv3x3_f32out(xs, wk, conv.bias, y, relu=True)
            acts.append((x, y))
            x = y
        feats.append(x)
        pool = (2, 2, 2, 2)
    partials, grads = [], []
    for k, f in enumerate(feats):
        lin_w = getattr(mod, f"lin{k}").model[1].weight
        p_, g_ = ops.lpips_head_f32(f, lin_w.detach().reshape(-1).contiguous(), want_grad)
        partials.append(p_)
        grads.append(g_)
    val = ops.lpips_finalize(partials, [f.shape[1] * f.shape[2] for f in feats], N)
    saved = (acts, feats, grads, N, normalize) if want_grad else None
    return val, saved


def _backward_impl_f32(mod, saved, gval):
    acts, feats, grads, N, normalize = saved
    pk = mod._packed
    dev = gval.device
    g = grads[4]                  # d / d relu5_3, already masked by the feature's own ReLU
    mask = None                   # ReLU mask still to be applied to g (folded into its split)
    ci = len(acts)
    for k in range(4, -1, -1):
        convs = _SLICES[k]
        for j in range(len(convs) - 1, -1, -1):
            idx, cin, cout = convs[j]
            ci -= 1
            conv = getattr(getattr(mod.net, f"slice{k + 1}"), str(idx))
            cbo = _pad16(cout)
            wd = pk.get((k, idx, "d3"), conv.weight, lambda t: ops.repack_conv3x3(ops.split_weights(t, 0, 3, cbo), True))
            ds = ops.split_stack(g, 3, cbo, relu_mask=mask)
            xin = acts[ci][0]
            dx = torch.empty((N, ds.shape[1], ds.shape[2], _pad16(cin)), dtype=torch.float32, device=dev)
            ops.conv3x3_f32out(ds, wd, None, dx)
            g = dx
            mask = xin[:N] if j > 0 else None          # ReLU of the producing convolution (same slice)
        if k > 0:
            f = feats[k - 1][:N]                       # g is the gradient w.r.t. the pooled tensor
            g = ops.maxpool_f32_bwd(f, g, torch.empty(f.shape, dtype=torch.float32, device=dev), 2, gskip=grads[k - 1])
            mask = f                                   # relu{k}'s mask, applied when g is split for the next dgrad
    return ops.lpips_unpack_grad_f32(g, gval.contiguous(), mod._shift_host, mod._scale_host, normalize)


class _LPIPSFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, in0, in1, normalize):
        fwd = _forward_impl_f32 if mod.precision == "fp32x" else _forward_impl
        val, saved = fwd(mod, in0, in1, normalize, ctx.needs_input_grad[1])
        ctx.mod, ctx.saved = mod, saved
        return val.view(-1, 1, 1, 1)

    @staticmethod
    def backward(ctx, gval):
        if ctx.saved is None:
            return None, None, None, None
        bwd = _backward_impl_f32 if ctx.mod.precision == "fp32x" else _backward_impl
        gin0 = bwd(ctx.mod, ctx.saved, gval.reshape(-1).float())
        ctx.saved = None
        return None, gin0, None, None


class ScalingLayer(nn.Module):
    def __init__(self):
        super().__init__()
        self.register_buffer("shift", torch.Tensor([-.030, -.088, -.188])[None, :, None, None])
        self.register_buffer("scale", torch.Tensor([.458, .448, .450])[None, :, None, None])


class NetLinLayer(nn.Module):
    """A single linear layer which does a 1x1 conv (lpips/lpips.py): Dropout (inactive: the package runs in
    eval mode) + Conv2d(chn_in, 1, 1, bias=False)."""

    def __init__(self, chn_in, chn_out=1, use_dropout=False):
        super().__init__()
        layers = [nn.Dropout()] if use_dropout else [nn.Identity()]
        layers += [nn.Conv2d(chn_in, chn_out, 1, stride=1, padding=0, bias=False)]
        self.model = nn.Sequential(*layers)


class _VGG16Slices(nn.Module):
    """Parameter container with the key layout of lpips.pretrained_networks.vgg16: slice1..slice5 hold the
    torchvision `features` layers under their original indices."""

    def __init__(self):
        super().__init__()
        for k, convs in enumerate(_SLICES):
            seq = nn.Sequential()
            for idx, cin, cout in convs:
                seq.add_module(str(idx), nn.Conv2d(cin, cout, kernel_size=3, padding=1))
            setattr(self, f"slice{k + 1}", seq)
        for p in self.parameters():
            p.requires_grad = False


class LPIPS(nn.Module):
    """lpips.LPIPS(net='vgg') (v0.1): forward(in0, in1, retPerLayer=False, normalize=False) -> [N, 1, 1, 1]."""

    def __init__(self, pretrained=False, net="vgg", version="0.1", lpips=True, spatial=False, pnet_rand=False,
                 pnet_tune=False, use_dropout=True, model_path=None, eval_mode=True, verbose=False, precision="bf16"):
        """precision (not in the package): "bf16" = bf16 tensor-core convolutions (fast; values 3e-4, parameter-level
        gradients 2e-2 from fp32), "fp32x" = emulated fp32 on the tensor cores (pixel gradient also fp32-class)."""
        super().__init__()
        if precision not in ("bf16", "fp32x"):
            raise ValueError(precision)
        self.precision = precision
        if net not in ("vgg", "vgg16") or not lpips or spatial or pnet_tune:
            raise NotImplementedError("only LPIPS(net='vgg', lpips=True, spatial=False) — the configuration the "
                                      "reference uses (rovr/train_local_net_unet.py:91, rovr/rovr.py:54)")
        if pretrained and model_path is None:
            raise RuntimeError("pretrained LPIPS/VGG weights need a download; there is no network here. Construct with "
                               "pretrained=False and load_state_dict() the package's weights (same key layout)")
        self.scaling_layer = ScalingLayer()
        self.net = _VGG16Slices()
        self.chns = list(_CHNS)
        self.L = len(self.chns)
        for k, c in enumerate(self.chns):
            lin = NetLinLayer(c, use_dropout=use_dropout)
            with torch.no_grad():       # the trained lin weights are non-negative; keep random ones so as well
                lin.model[1].weight.abs_().mul_(4.0)
            lin.model[1].weight.requires_grad = False
            setattr(self, f"lin{k}", lin)
        self._packed = PackedWeights()
        self._shift_host = [-.030, -.088, -.188]
        self._scale_host = [.458, .448, .450]
        if model_path is not None:
            self.load_state_dict(torch.load(model_path, map_location="cpu"), strict=False)
        if eval_mode:
            self.eval()

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        # the package also registers the lin layers as `lins` (ModuleList of the same modules): drop the aliases
        for k in [k for k in state_dict if k.startswith(prefix + "lins.")]:
            state_dict.pop(k)
        # host copies of the ScalingLayer constants (kernel arguments): follow a loaded checkpoint
        for name, attr in (("shift", "_shift_host"), ("scale", "_scale_host")):
            t = state_dict.get(prefix + "scaling_layer." + name)
            if t is not None:
                setattr(self, attr, [float(v) for v in t.detach().cpu().flatten().tolist()])
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)

    def forward(self, in0, in1, retPerLayer=False, normalize=False):
        if retPerLayer:
            raise NotImplementedError("retPerLayer is not used by the reference")
        if not in0.is_cuda:
            raise RuntimeError("LPIPS (B200) needs CUDA tensors: there is no CPU path")
        if in0.shape != in1.shape or in0.dim() != 4 or in0.shape[1] != 3:
            raise ValueError(f"expected two [N, 3, H, W] batches, got {tuple(in0.shape)} and {tuple(in1.shape)}")
        if in0.shape[2] % 16 or in0.shape[3] % 16:
            raise ValueError("H and W must be multiples of 16 (four 2x2 poolings)")
        if in1.requires_grad and torch.is_grad_enabled():
            raise NotImplementedError("gradient w.r.t. the second image (the target) is not implemented")
        with torch.cuda.device(in0.device):
            return _LPIPSFunction.apply(self, in0.float().contiguous(), in1.detach().float().contiguous(), bool(normalize))
