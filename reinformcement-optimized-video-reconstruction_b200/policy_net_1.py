"""Drop-in for the reference `policy_net_1` module: `PolicyNetwork1UNet(is_critic=False)`.

Same class name, constructor argument, methods (`unet`, `compute_logits`, `forward(image, context,
device=None)`, `logprob(image, context, action)`), attribute names, parameter order and
state_dict keys as rovr/policy_net_1.py:10-115. The half-width U-Net with train-mode BatchNorm
after every convolution (:60-84) runs inside one autograd.Function on the B200 kernels:

    pack(cat[image, context]) -> NHWC bf16 (6 -> 16 channels)        (:88)
    conv1..conv4 / conv5..conv7   tcgen05 implicit GEMM, bias in the epilogue
    bn* + ReLU                    fused statistics + normalise kernels (batch statistics, running
                                  stats updated like nn.BatchNorm2d), written straight into the
                                  [up, skip] concat buffers (:69,73,77)
    upconv1..3                    GEMM + pixel-shuffle store
    conv8 / conv9 (1x1)           GEMM with channels padded to 16
    2x2 max-pools                 pool kernels

and the head (per-sample standardisation with unbiased std and no eps, fc_final, gumbel-softmax,
max / gather) runs in fp32 head kernels. The actor's `forward` is only shape-valid for b == 1 (or
b == 25) in the reference (`logits - logits.mean(dim=1)` without keepdim, :100); the same rule is
enforced here.
"""
import torch
import torch.nn as nn

import ops
import os

from _blocks import BF, F32, PackedWeights, TrunkOps, TrunkOpsF32
from _heads import GumbelLogProb, LinearF32, MaskedLogits, Standardize, exponential_like

_CONV = ["conv1", "conv2", "conv3", "conv4", "upconv1", "conv5", "upconv2", "conv6", "upconv3", "conv7",
         "conv8", "conv9"]
_BN = {"conv1": "bn1", "conv2": "bn2", "conv3": "bn3", "conv4": "bn4", "upconv1": "bn_up1", "conv5": "bn5",
       "upconv2": "bn_up2", "conv6": "bn6", "upconv3": "bn_up3", "conv7": "bn7", "conv8": "bn8", "conv9": "bn9"}
_LIVE = [n + s for c in _CONV for n in (c, _BN[c]) for s in (".weight", ".bias")]


class _UNetFeatures(torch.autograd.Function):
    """inp [b,6,H,W] fp32 -> unet(inp) [b,1,H/4,W/4] fp32 (rovr/policy_net_1.py:60-84)."""

    @staticmethod
    def forward(ctx, net, image, context, *plist):
        P = dict(zip(_LIVE, plist))
        T = TrunkOps(P, dict(net.named_buffers()), net._packed, training=net.training)
        dev = image.device
        b, _, H, W = image.shape

        def buf(div, c):
            return torch.empty((b, H // div, W // div, c), dtype=BF, device=dev)

        a = {}
        a["in16"] = ops.pack_nchw([image, context], 16)
        a["cat7"], a["cat6"], a["cat5"] = buf(1, 64), buf(2, 128), buf(4, 256)
        x1, x2, x3 = a["cat7"][..., 32:], a["cat6"][..., 64:], a["cat5"][..., 128:]
        a["r1"] = T.cbr3_fwd("conv1", "bn1", a["in16"], x1)
        a["p1"] = ops.maxpool_fwd(x1, buf(2, 32), 2)
        a["r2"] = T.cbr3_fwd("conv2", "bn2", a["p1"], x2)
        a["p2"] = ops.maxpool_fwd(x2, buf(4, 64), 2)
        a["r3"] = T.cbr3_fwd("conv3", "bn3", a["p2"], x3)
        a["p3"] = ops.maxpool_fwd(x3, buf(8, 128), 2)
        a["x4"] = buf(8, 256)
        a["r4"] = T.cbr3_fwd("conv4", "bn4", a["p3"], a["x4"])

        a["ru1"] = T.ubr_fwd("upconv1", "bn_up1", a["x4"], a["cat5"][..., :128])
        a["y5"] = buf(4, 128)
        a["r5"] = T.cbr3_fwd("conv5", "bn5", a["cat5"], a["y5"])
        a["ru2"] = T.ubr_fwd("upconv2", "bn_up2", a["y5"], a["cat6"][..., :64])
        a["y6"] = buf(2, 64)
        a["r6"] = T.cbr3_fwd("conv6", "bn6", a["cat6"], a["y6"])
        a["ru3"] = T.ubr_fwd("upconv3", "bn_up3", a["y6"], a["cat7"][..., :32])
        a["y7"] = buf(1, 32)
        a["r7"] = T.cbr3_fwd("conv7", "bn7", a["cat7"], a["y7"])

        a["y8"] = buf(1, 16)                                        # 3 valid channels
        a["r8"] = T.cbr1_fwd("conv8", "bn8", a["y7"], a["y8"])
        a["p8"] = ops.maxpool_fwd(a["y8"], buf(2, 16), 2)
        a["y9"] = buf(2, 16)                                        # 1 valid channel
        a["r9"] = T.cbr1_fwd("conv9", "bn9", a["p8"], a["y9"])
        a["p9"] = ops.maxpool_fwd(a["y9"], buf(4, 16), 2)
        out = torch.empty((b, (H // 4) * (W // 4)), dtype=torch.float32, device=dev)
        ops.flatten_nhwc(a["p9"], 1, out, 0)
        ctx.acts, ctx.T = a, T
        return out.view(b, 1, H // 4, W // 4)

    @staticmethod
    def backward(ctx, g):
        a, T = ctx.acts, ctx.T
        b = g.shape[0]
        g = g.contiguous().float().view(b, -1)

        def like(t):
            return torch.empty_like(t)

        x1, x2, x3 = a["cat7"][..., 32:], a["cat6"][..., 64:], a["cat5"][..., 128:]
        gp9 = ops.unflatten_nhwc(g, 1, like(a["p9"]), 0)
        gy9 = ops.maxpool_bwd(a["y9"], gp9, like(a["y9"]), 2, relu_mask=False)
        gp8 = like(a["p8"])
        T.cbr1_bwd("conv9", "bn9", a["r9"], a["p8"], a["y9"], gy9, gp8)
        gy8 = ops.maxpool_bwd(a["y8"], gp8, like(a["y8"]), 2, relu_mask=False)
        gy7 = like(a["y7"])
        T.cbr1_bwd("conv8", "bn8", a["r8"], a["y7"], a["y8"], gy8, gy7)
        gcat7 = like(a["cat7"])
        T.cbr3_bwd("conv7", "bn7", a["r7"], a["cat7"], a["y7"], gy7, gcat7)
        gy6 = like(a["y6"])
        T.ubr_bwd("upconv3", "bn_up3", a["ru3"], a["y6"], a["cat7"][..., :32], gcat7[..., :32], gy6)
        gcat6 = like(a["cat6"])
        T.cbr3_bwd("conv6", "bn6", a["r6"], a["cat6"], a["y6"], gy6, gcat6)
        gy5 = like(a["y5"])
        T.ubr_bwd("upconv2", "bn_up2", a["ru2"], a["y5"], a["cat6"][..., :64], gcat6[..., :64], gy5)
        gcat5 = like(a["cat5"])
        T.cbr3_bwd("conv5", "bn5", a["r5"], a["cat5"], a["y5"], gy5, gcat5)
        gx4 = like(a["x4"])
        T.ubr_bwd("upconv1", "bn_up1", a["ru1"], a["x4"], a["cat5"][..., :128], gcat5[..., :128], gx4)
        gp3 = like(a["p3"])
        T.cbr3_bwd("conv4", "bn4", a["r4"], a["p3"], a["x4"], gx4, gp3)
        # skip connections: d x_k = (gradient through the concat) + (gradient through the pool)
        gx3 = torch.empty(x3.shape, dtype=BF, device=g.device)
        ops.maxpool_bwd(x3, gp3, gx3, 2, gskip=gcat5[..., 128:], relu_mask=False)
        gp2 = like(a["p2"])
        T.cbr3_bwd("conv3", "bn3", a["r3"], a["p2"], x3, gx3, gp2)
        gx2 = torch.empty(x2.shape, dtype=BF, device=g.device)
        ops.maxpool_bwd(x2, gp2, gx2, 2, gskip=gcat6[..., 64:], relu_mask=False)
        gp1 = like(a["p1"])
        T.cbr3_bwd("conv2", "bn2", a["r2"], a["p1"], x2, gx2, gp1)
        gx1 = torch.empty(x1.shape, dtype=BF, device=g.device)
        ops.maxpool_bwd(x1, gp1, gx1, 2, gskip=gcat7[..., 32:], relu_mask=False)
        T.cbr3_bwd("conv1", "bn1", a["r1"], a["in16"], x1, gx1, None)   # inputs need no gradient
        ctx.acts = None
        return (None, None, None) + tuple(T.G[n] for n in _LIVE)


class _UNetFeaturesF32(torch.autograd.Function):
    """The same network on the emulated-fp32 path (fp32 NHWC activations, split-bf16 stacked tensor-core
    products, _blocks.TrunkOpsF32): outputs and gradients match the fp32 reference to ~1e-4, which the
    bf16-operand path cannot (every gradient passes the two 2x2 max-pools of :80-82)."""

    @staticmethod
    def forward(ctx, net, image, context, *plist):
        P = dict(zip(_LIVE, plist))
        T = TrunkOpsF32(P, dict(net.named_buffers()), net._packed, training=net.training)
        dev = image.device
        b, _, H, W = image.shape
        P2 = (2, 2, 2, 2)

        def buf(div, c):
            return torch.empty((b, H // div, W // div, c), dtype=F32, device=dev)

        a = {}
        xs = ops.split_stack(image, 6, 8, layout="nchw", src2=context)      # cat([image, context], 1) (:88)
        a["cat7"], a["cat6"], a["cat5"] = buf(1, 64), buf(2, 128), buf(4, 256)
        x1, x2, x3 = a["cat7"][..., 32:], a["cat6"][..., 64:], a["cat5"][..., 128:]
        a["r1"] = T.cbr3_fwd("conv1", "bn1", xs, x1)
        a["r2"] = T.cbr3_fwd("conv2", "bn2", T.stack6(x1, P2), x2)           # MaxPool2d(2, 2) fused into the split
        a["r3"] = T.cbr3_fwd("conv3", "bn3", T.stack6(x2, P2), x3)
        a["x4"] = buf(8, 256)
        a["r4"] = T.cbr3_fwd("conv4", "bn4", T.stack6(x3, P2), a["x4"])
        a["ru1"] = T.ubr_fwd("upconv1", "bn_up1", T.stack6(a["x4"]), a["cat5"][..., :128])
        a["y5"] = buf(4, 128)
        a["r5"] = T.cbr3_fwd("conv5", "bn5", T.stack6(a["cat5"]), a["y5"])
        a["ru2"] = T.ubr_fwd("upconv2", "bn_up2", T.stack6(a["y5"]), a["cat6"][..., :64])
        a["y6"] = buf(2, 64)
        a["r6"] = T.cbr3_fwd("conv6", "bn6", T.stack6(a["cat6"]), a["y6"])
        a["ru3"] = T.ubr_fwd("upconv3", "bn_up3", T.stack6(a["y6"]), a["cat7"][..., :32])
        a["y7"] = buf(1, 32)
        a["r7"] = T.cbr3_fwd("conv7", "bn7", T.stack6(a["cat7"]), a["y7"])
        a["y8"] = buf(1, 16)                                                 # 3 valid channels
        a["r8"] = T.cbr1_fwd("conv8", "bn8", T.stack6(a["y7"]), a["y8"])
        a["y9"] = buf(2, 16)                                                 # 1 valid channel
        a["r9"] = T.cbr1_fwd("conv9", "bn9", T.stack6(a["y8"][..., :8], P2), a["y9"])
        a["p9"] = ops.maxpool_f32_fwd(a["y9"], 2)
        out = torch.empty((b, (H // 4) * (W // 4)), dtype=F32, device=dev)
        ops.flatten_f32(a["p9"], 1, out, 0)
        ctx.acts, ctx.T = a, T
        return out.view(b, 1, H // 4, W // 4)

    @staticmethod
    def backward(ctx, g):
        a, T = ctx.acts, ctx.T
        b = g.shape[0]
        g = g.contiguous().float().view(b, -1)

        def like(t):
            return torch.empty(t.shape, dtype=F32, device=t.device)

        x1, x2, x3 = a["cat7"][..., 32:], a["cat6"][..., 64:], a["cat5"][..., 128:]
        gp9 = ops.unflatten_f32(g, 1, like(a["p9"]), 0)
        gy9 = ops.maxpool_f32_bwd(a["y9"], gp9, like(a["y9"]), 2)
        gp8 = torch.empty(a["y9"].shape, dtype=F32, device=g.device)         # pooled y8: same pixels as y9
        T.cbr1_bwd("conv9", "bn9", a["r9"], a["y9"], gy9, gp8)
        gy8 = ops.maxpool_f32_bwd(a["y8"], gp8, like(a["y8"]), 2)
        gy7 = like(a["y7"])
        T.cbr1_bwd("conv8", "bn8", a["r8"], a["y8"], gy8, gy7)
        gcat7 = like(a["cat7"])
        T.cbr3_bwd("conv7", "bn7", a["r7"], a["y7"], gy7, gcat7)
        gy6 = like(a["y6"])
        T.ubr_bwd("upconv3", "bn_up3", a["ru3"], a["cat7"][..., :32], gcat7[..., :32], gy6)
        gcat6 = like(a["cat6"])
        T.cbr3_bwd("conv6", "bn6", a["r6"], a["y6"], gy6, gcat6)
        gy5 = like(a["y5"])
        T.ubr_bwd("upconv2", "bn_up2", a["ru2"], a["cat6"][..., :64], gcat6[..., :64], gy5)
        gcat5 = like(a["cat5"])
        T.cbr3_bwd("conv5", "bn5", a["r5"], a["y5"], gy5, gcat5)
        gx4 = like(a["x4"])
        T.ubr_bwd("upconv1", "bn_up1", a["ru1"], a["cat5"][..., :128], gcat5[..., :128], gx4)
        H8, W8 = a["x4"].shape[1], a["x4"].shape[2]
        gp3 = torch.empty((b, H8, W8, 128), dtype=F32, device=g.device)
        T.cbr3_bwd("conv4", "bn4", a["r4"], a["x4"], gx4, gp3)
        # skip connections: d x_k = (gradient through the concat) + (gradient through the pool)
        gx3 = ops.maxpool_f32_bwd(x3, gp3, like(x3), 2, gskip=gcat5[..., 128:])
        gp2 = torch.empty((b, 2 * H8, 2 * W8, 64), dtype=F32, device=g.device)
        T.cbr3_bwd("conv3", "bn3", a["r3"], x3, gx3, gp2)
        gx2 = ops.maxpool_f32_bwd(x2, gp2, like(x2), 2, gskip=gcat6[..., 64:])
        gp1 = torch.empty((b, 4 * H8, 4 * W8, 32), dtype=F32, device=g.device)
        T.cbr3_bwd("conv2", "bn2", a["r2"], x2, gx2, gp1)
        gx1 = ops.maxpool_f32_bwd(x1, gp1, like(x1), 2, gskip=gcat7[..., 32:])
        T.cbr3_bwd("conv1", "bn1", a["r1"], x1, gx1, None)                  # inputs need no gradient
        ctx.acts = None
        return (None, None, None) + tuple(T.G[n] for n in _LIVE)


def default_trunk_precision():
    """"fp32x" (emulated fp32 on the tensor cores; parity with the fp32 reference) unless
    ROVR_POLICY_PRECISION=bf16 asks for the plain bf16-operand trunks (half the memory, gradients behind
    the max-pools then deviate from the fp32 reference — DESIGN.md §6)."""
    v = os.environ.get("ROVR_POLICY_PRECISION", "fp32x")
    if v not in ("fp32x", "bf16"):
        raise ValueError(f"ROVR_POLICY_PRECISION must be fp32x or bf16, got {v!r}")
    return v


class PolicyNetwork1UNet(nn.Module):
    """Reference: rovr/policy_net_1.py:10-115."""

    def __init__(self, is_critic=False):
        super(PolicyNetwork1UNet, self).__init__()
        self.num_composed_frames = 25
        self.num_channels = 3
        self.is_critic = is_critic
        self.temperature = .5
        self.image_size = 80

        self.conv1 = nn.Conv2d(6, 32, kernel_size=3, padding=1)
        self.bn1 = nn.BatchNorm2d(32)
        self.conv2 = nn.Conv2d(32, 64, kernel_size=3, padding=1)
        self.bn2 = nn.BatchNorm2d(64)
        self.conv3 = nn.Conv2d(64, 128, kernel_size=3, padding=1)
        self.bn3 = nn.BatchNorm2d(128)
        self.conv4 = nn.Conv2d(128, 256, kernel_size=3, padding=1)
        self.bn4 = nn.BatchNorm2d(256)
        self.maxpool = nn.MaxPool2d(kernel_size=2, stride=2)
        self.upconv1 = nn.ConvTranspose2d(256, 128, kernel_size=2, stride=2)
        self.bn_up1 = nn.BatchNorm2d(128)
        self.conv5 = nn.Conv2d(256, 128, kernel_size=3, padding=1)
        self.bn5 = nn.BatchNorm2d(128)
        self.upconv2 = nn.ConvTranspose2d(128, 64, kernel_size=2, stride=2)
        self.bn_up2 = nn.BatchNorm2d(64)
        self.conv6 = nn.Conv2d(128, 64, kernel_size=3, padding=1)
        self.bn6 = nn.BatchNorm2d(64)
        self.upconv3 = nn.ConvTranspose2d(64, 32, kernel_size=2, stride=2)
        self.bn_up3 = nn.BatchNorm2d(32)
        self.conv7 = nn.Conv2d(64, 32, kernel_size=3, padding=1)
        self.bn7 = nn.BatchNorm2d(32)
        self.conv8 = nn.Conv2d(32, 3, kernel_size=1)
        self.bn8 = nn.BatchNorm2d(3)
        self.conv9 = nn.Conv2d(3, 1, kernel_size=1)
        self.bn9 = nn.BatchNorm2d(1)
        self.dropout = 0.1
        self.fc_final = nn.Linear(400, 1 if self.is_critic else 25)
        self._packed = PackedWeights()
        self.trunk_precision = default_trunk_precision()

    def _features(self, image, context):
        if not image.is_cuda:
            raise RuntimeError("PolicyNetwork1UNet (B200) needs CUDA tensors: there is no CPU path")
        if image.shape[2] % 8 or image.shape[3] % 8:
            raise ValueError("H and W must be multiples of 8 (three 2x2 poolings)")
        named = dict(self.named_parameters())
        fn = _UNetFeaturesF32 if self.trunk_precision == "fp32x" else _UNetFeatures
        with torch.cuda.device(image.device):
            return fn.apply(self, image.float().contiguous(), context.float().contiguous(),
                            *[named[n] for n in _LIVE])

    def unet(self, x):
        """x: [b, 6, h, w] (rovr/policy_net_1.py:60-84)."""
        return self._features(x[:, :3], x[:, 3:])

    def compute_logits(self, x, context):
        image = self._features(x, context)                          # cat([x, context], 1) is the pack (:88)
        image = image.flatten(1)                                    # 'b c h w -> b (c h w)' (:90)
        normalized_image = Standardize.apply(image, 1, 0.0)         # unbiased std, no eps (:91-93)
        return LinearF32.apply(normalized_image, self.fc_final.weight, self.fc_final.bias)

    def forward(self, image, context, device=None):
        logits = self.compute_logits(image, context)
        if not self.is_critic:
            std_logits = MaskedLogits.apply(logits.detach(), None, True)   # (:100), b == 1 or 25 only
            expo = exponential_like(std_logits)
            _, idx, logp = ops.head_gumbel_fwd(std_logits, expo, self.temperature, 1)
            return idx, logp                                        # both detached (:103)
        return logits.squeeze(1)

    def logprob(self, image, context, action):
        if self.is_critic:
            raise Exception("DO NOT CALL LOGPROB FOR CRITIC")
        logits = self.compute_logits(image, context)
        expo = exponential_like(logits)
        return GumbelLogProb.apply(logits, expo, self.temperature, 3, action.to(logits.device))
