"""Drop-in for the reference `policy_net_2` module: `PolicyNetwork2UNet(is_critic=False)`.

Same class name, constructor argument, method signatures (`forward(image, context, target,
device=None, extra=None)`, `compute_logits`, `get_masked_logits`, `logprob`), attribute names,
parameter order and state_dict keys as rovr/policy_net_2.py:10-141 — including the unused
`context_conv` branch (:27-38, 1.48 M dead parameters that receive no gradient) — but

  * `video_conv` (:41-60; 4 x [Conv3x3 + train-mode BatchNorm + ReLU + MaxPool]) runs as tcgen05
    implicit-GEMM convolutions + fused BN/ReLU/pool kernels inside one autograd.Function;
  * `final_fc` (:63-69; five activation-free Linears) runs as fp32 weight-streaming kernels;
  * scatter / standardisation quirks / gumbel-softmax / top-2 run as fp32 head kernels, so the
    selected frame indices are a deterministic function of the logits and torch's RNG draw.

The reference's debugging print() calls (:78,83,90,94,114,119,135) are not reproduced.
"""
import torch
import torch.nn as nn

import ops
from _blocks import BF, F32, PackedWeights, TrunkOps, TrunkOpsF32
from _heads import GumbelLogProb, LinearF32, MaskedLogits, Standardize, exponential_like

_VC = [("video_conv.0", "video_conv.1"), ("video_conv.4", "video_conv.5"),
       ("video_conv.8", "video_conv.9"), ("video_conv.12", "video_conv.13")]
_LIVE = [n + s for pair in _VC for n in pair for s in (".weight", ".bias")]


class _VideoConvStacked(torch.autograd.Function):
    """(image [b,1,160,160], feature rows [b,F]) -> cat([video_conv(image), features], 1)."""

    @staticmethod
    def forward(ctx, net, image, feat, *plist):
        P = dict(zip(_LIVE, plist))
        T = TrunkOps(P, dict(net.named_buffers()), net._packed, training=net.training)
        dev = image.device
        b, _, H, W = image.shape

        def buf(h, w, c):
            return torch.empty((b, h, w, c), dtype=BF, device=dev)

        a = {}
        a["in16"] = ops.pack_nchw([image], 16)
        a["a0"] = buf(H, W, 64)
        a["r0"] = T.cbr3_fwd(*_VC[0], a["in16"], a["a0"])
        a["p0"] = ops.maxpool_fwd(a["a0"], buf(H // 8, W // 8, 64), 8)
        h1, w1 = H // 8, W // 8
        a["a4"] = buf(h1, w1, 128)
        a["r4"] = T.cbr3_fwd(*_VC[1], a["p0"], a["a4"])
        h2, w2 = h1 // 4, w1 // 4
        a["p4"] = ops.maxpool_fwd(a["a4"], buf(h2, w2, 128), 4)
        a["a8"] = buf(h2, w2, 256)                       # MaxPool2d(1, 1) is the identity (:53)
        a["r8"] = T.cbr3_fwd(*_VC[2], a["p4"], a["a8"])
        a["a12"] = buf(h2, w2, 512)
        a["r12"] = T.cbr3_fwd(*_VC[3], a["a8"], a["a12"])
        h3, w3 = (h2 - 2) // 2 + 1, (w2 - 2) // 1 + 1    # MaxPool2d(2, stride=(2, 1))  (:57)
        a["q"] = ops.maxpool_fwd(a["a12"], buf(h3, w3, 512), 2, (2, 1))
        h4, w4 = (h3 - 2) // 2 + 1, (w3 - 2) // 2 + 1    # MaxPool2d(2, stride=(2, 2))  (:58)
        a["r"] = ops.maxpool_fwd(a["q"], buf(h4, w4, 512), 2, (2, 2))
        nvid = 512 * h4 * w4
        stacked = torch.empty((b, nvid + feat.shape[1]), dtype=torch.float32, device=dev)
        ops.flatten_nhwc(a["r"], 512, stacked, 0)        # nn.Flatten on NCHW (:59)
        ops.copy2d_f32(feat, stacked[:, nvid:])          # torch.cat([vector_out, image_out], 1) (:92)
        ctx.net, ctx.acts, ctx.T, ctx.nvid = net, a, T, nvid
        ctx.feat_needs_grad = feat.requires_grad
        return stacked

    @staticmethod
    def backward(ctx, g):
        a, T, nvid = ctx.acts, ctx.T, ctx.nvid
        g = g.contiguous().float()

        def like(t):
            return torch.empty_like(t)

        gr = ops.unflatten_nhwc(g, 512, like(a["r"]), 0)
        gq = ops.maxpool_bwd(a["q"], gr, like(a["q"]), 2, (2, 2), relu_mask=False)
        ga12 = ops.maxpool_bwd(a["a12"], gq, like(a["a12"]), 2, (2, 1), relu_mask=False)
        ga8 = like(a["a8"])
        T.cbr3_bwd(*_VC[3], a["r12"], a["a8"], a["a12"], ga12, ga8)
        gp4 = like(a["p4"])
        T.cbr3_bwd(*_VC[2], a["r8"], a["p4"], a["a8"], ga8, gp4)
        ga4 = ops.maxpool_bwd(a["a4"], gp4, like(a["a4"]), 4, relu_mask=False)
        gp0 = like(a["p0"])
        T.cbr3_bwd(*_VC[1], a["r4"], a["p0"], a["a4"], ga4, gp0)
        ga0 = ops.maxpool_bwd(a["a0"], gp0, like(a["a0"]), 8, relu_mask=False)
        T.cbr3_bwd(*_VC[0], a["r0"], a["in16"], a["a0"], ga0, None)   # the mosaic needs no gradient
        gfeat = None
        if ctx.feat_needs_grad:
            gfeat = torch.empty((g.shape[0], g.shape[1] - nvid), dtype=torch.float32, device=g.device)
            ops.copy2d_f32(g[:, nvid:], gfeat)
        ctx.acts = None
        return (None, None, gfeat) + tuple(T.G[n] for n in _LIVE)


class _VideoConvStackedF32(torch.autograd.Function):
    """The same on the emulated-fp32 path (_blocks.TrunkOpsF32): the 8x8 and 4x4 max-pools decide their
    arg-max on fp32-class convolution outputs, so the gradients match the fp32 reference."""

    @staticmethod
    def forward(ctx, net, image, feat, *plist):
        P = dict(zip(_LIVE, plist))
        T = TrunkOpsF32(P, dict(net.named_buffers()), net._packed, training=net.training)
        dev = image.device
        b, _, H, W = image.shape

        def buf(h, w, c):
            return torch.empty((b, h, w, c), dtype=F32, device=dev)

        a = {}
        xs = ops.split_stack(image, 6, 8, layout="nchw")                 # [b, H, W, 48]: 1 valid channel per block
        a["a0"] = buf(H, W, 64)
        a["r0"] = T.cbr3_fwd(*_VC[0], xs, a["a0"])
        h1, w1 = H // 8, W // 8
        a["a4"] = buf(h1, w1, 128)
        a["r4"] = T.cbr3_fwd(*_VC[1], T.stack6(a["a0"], (8, 8, 8, 8)), a["a4"])      # MaxPool2d(8, 8) (:46)
        h2, w2 = h1 // 4, w1 // 4
        a["a8"] = buf(h2, w2, 256)
        a["r8"] = T.cbr3_fwd(*_VC[2], T.stack6(a["a4"], (4, 4, 4, 4)), a["a8"])      # MaxPool2d(4, 4) (:50)
        a["a12"] = buf(h2, w2, 512)                                      # MaxPool2d(1, 1) is the identity (:53)
        a["r12"] = T.cbr3_fwd(*_VC[3], T.stack6(a["a8"]), a["a12"])
        a["q"] = ops.maxpool_f32_fwd(a["a12"], 2, (2, 1))                # :57
        a["r"] = ops.maxpool_f32_fwd(a["q"], 2, (2, 2))                  # :58
        h4, w4 = a["r"].shape[1], a["r"].shape[2]
        nvid = 512 * h4 * w4
        stacked = torch.empty((b, nvid + feat.shape[1]), dtype=F32, device=dev)
        ops.flatten_f32(a["r"], 512, stacked, 0)                         # nn.Flatten on NCHW (:59)
        ops.copy2d_f32(feat, stacked[:, nvid:])                          # torch.cat([vector_out, image_out], 1) (:92)
        ctx.net, ctx.acts, ctx.T, ctx.nvid = net, a, T, nvid
        ctx.feat_needs_grad = feat.requires_grad
        ctx.dims = (h1, w1)
        return stacked

    @staticmethod
    def backward(ctx, g):
        a, T, nvid = ctx.acts, ctx.T, ctx.nvid
        g = g.contiguous().float()
        b = g.shape[0]
        h1, w1 = ctx.dims

        def like(t):
            return torch.empty(t.shape, dtype=F32, device=t.device)

        gr = ops.unflatten_f32(g, 512, like(a["r"]), 0)
        gq = ops.maxpool_f32_bwd(a["q"], gr, like(a["q"]), 2, (2, 2))
        ga12 = ops.maxpool_f32_bwd(a["a12"], gq, like(a["a12"]), 2, (2, 1))
        ga8 = like(a["a8"])
        T.cbr3_bwd(*_VC[3], a["r12"], a["a12"], ga12, ga8)
        gp4 = torch.empty((b, a["a8"].shape[1], a["a8"].shape[2], 128), dtype=F32, device=g.device)
        T.cbr3_bwd(*_VC[2], a["r8"], a["a8"], ga8, gp4)
        ga4 = ops.maxpool_f32_bwd(a["a4"], gp4, like(a["a4"]), 4)
        gp0 = torch.empty((b, h1, w1, 64), dtype=F32, device=g.device)
        T.cbr3_bwd(*_VC[1], a["r4"], a["a4"], ga4, gp0)
        ga0 = ops.maxpool_f32_bwd(a["a0"], gp0, like(a["a0"]), 8)
        T.cbr3_bwd(*_VC[0], a["r0"], a["a0"], ga0, None)                 # the mosaic needs no gradient
        gfeat = None
        if ctx.feat_needs_grad:
            gfeat = torch.empty((g.shape[0], g.shape[1] - nvid), dtype=F32, device=g.device)
            ops.copy2d_f32(g[:, nvid:], gfeat)
        ctx.acts = None
        return (None, None, gfeat) + tuple(T.G[n] for n in _LIVE)


class PolicyNetwork2UNet(nn.Module):
    """Reference: rovr/policy_net_2.py:10-141."""

    def __init__(self, is_critic=False):
        super(PolicyNetwork2UNet, self).__init__()
        self.num_composed_frames = 20
        self.is_critic = is_critic
        self.output_size = 1 if self.is_critic else self.num_composed_frames
        self.context_size = 256
        self.num_channels = 1
        self.temperature = .7
        self.num_resnet_features = 2048

        # constructed for state_dict / parameter-order parity; never applied (rovr/policy_net_2.py:89)
        self.context_conv = nn.Sequential(
            nn.Conv2d(3, 128, kernel_size=3, stride=1, padding=1), nn.ReLU(), nn.MaxPool2d(kernel_size=8, stride=8),
            nn.Conv2d(128, 256, kernel_size=3, stride=1, padding=1), nn.ReLU(), nn.MaxPool2d(kernel_size=4, stride=4),
            nn.Conv2d(256, 512, kernel_size=3, stride=1, padding=1), nn.ReLU(), nn.MaxPool2d(kernel_size=4, stride=4),
            nn.Flatten())
        # parameter containers of the B200 trunk (same module layout as the reference)
        self.video_conv = nn.Sequential(
            nn.Conv2d(self.num_channels, 64, kernel_size=3, stride=1, padding=1), nn.BatchNorm2d(64), nn.ReLU(),
            nn.MaxPool2d(kernel_size=8, stride=8),
            nn.Conv2d(64, 128, kernel_size=3, stride=1, padding=1), nn.BatchNorm2d(128), nn.ReLU(),
            nn.MaxPool2d(kernel_size=4, stride=4),
            nn.Conv2d(128, 256, kernel_size=3, stride=1, padding=1), nn.BatchNorm2d(256), nn.ReLU(),
            nn.MaxPool2d(kernel_size=1, stride=1),
            nn.Conv2d(256, 512, kernel_size=3, stride=1, padding=1), nn.BatchNorm2d(512), nn.ReLU(),
            nn.MaxPool2d(kernel_size=2, stride=(2, 1)), nn.MaxPool2d(kernel_size=2, stride=(2, 2)),
            nn.Flatten())
        self.final_fc = nn.Sequential(
            nn.Linear(2048, 1024), nn.Linear(1024, 512), nn.Linear(512, 256), nn.Linear(256, 64),
            nn.Linear(64, self.output_size))
        self._packed = PackedWeights()
        from policy_net_1 import default_trunk_precision
        self.trunk_precision = default_trunk_precision()

    # -- trunk -----------------------------------------------------------------------------------
    def _stacked(self, image, context):
        if not image.is_cuda:
            raise RuntimeError("PolicyNetwork2UNet (B200) needs CUDA tensors: there is no CPU path")
        feat = context.squeeze(1)                                    # :90
        if image.dim() != 4 or image.shape[1] != 1 or feat.dim() != 2 or feat.shape[0] != image.shape[0]:
            raise ValueError(f"expected image [b,1,h,w] and context [b,1,f], got {tuple(image.shape)} "
                             f"{tuple(context.shape)}")
        named = dict(self.named_parameters())
        plist = [named[n] for n in _LIVE]
        fn = _VideoConvStackedF32 if self.trunk_precision == "fp32x" else _VideoConvStacked
        with torch.cuda.device(image.device):
            return fn.apply(self, image.float().contiguous(), feat.float().contiguous(), *plist)

    def compute_logits(self, x, device=None):
        """x: (b, 2048) -> final_fc(x)  (rovr/policy_net_2.py:71-79)."""
        for layer in self.final_fc:
            x = LinearF32.apply(x, layer.weight, layer.bias)
        return x

    def forward(self, image, context, target, device=None, extra=None):
        if self.is_critic:
            image = image.unsqueeze(1)
        stacked = self._stacked(image, context)
        if extra is not None:
            return self.get_masked_logits(stacked, target, device)
        if not self.is_critic:
            logits = self.get_masked_logits(stacked, target, device)
            expo = exponential_like(logits)
            _, idx, logprob = ops.head_gumbel_fwd(logits.detach(), expo, self.temperature, 2)
            return idx, logprob                                      # both detached (:102)
        stacked = Standardize.apply(stacked, 0, 0.001)               # over the BATCH dim (:104-106)
        return self.compute_logits(stacked, device).squeeze(1)

    def get_masked_logits(self, stacked, target, device=None):
        if self.is_critic:
            raise Exception("DO NOT CALL get_masked_logits FOR CRITIC")
        logits = self.compute_logits(stacked, device)
        target = target.to(torch.int64).squeeze(1)                   # :117
        if target.dim() != 2:
            raise RuntimeError("scatter_(): target must be [b, k] after squeeze(1), as in the reference "
                               f"(got {tuple(target.shape)})")
        return MaskedLogits.apply(logits, target.to(logits.device), True)

    def logprob(self, image, context, target, action, device):
        if self.is_critic:
            raise Exception("DO NOT CALL LOGPROB FOR CRITIC")
        image = image.unsqueeze(1)
        stacked = self._stacked(image, context)
        logits = self.compute_logits(stacked)
        logits = MaskedLogits.apply(logits, target.to(logits.device), False)     # scatter_(1, target, 0) (:138)
        expo = exponential_like(logits)
        return GumbelLogProb.apply(logits, expo, self.temperature, 4, action.to(logits.device))
