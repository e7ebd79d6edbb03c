"""Drop-in for the reference `resnet_extractor` module: `ResnetFeatureExtractor(pretrained=False)`.

Same class name, constructor argument, attributes (`resnet` = Sequential of torchvision
resnet50's children minus the FC, `linear` = Linear(2048, 768), `preprocessing`), methods
(`forward`, `calculate_index`, `encode`, `insert_encoded_frame_batch`, `extract_patch`) and
state_dict keys as rovr/resnet_extractor.py:5-67. torchvision is only the parameter container
(the reference builds the same object, :8,16); the arithmetic runs on the B200 kernels:

  * all frames of the call are encoded as ONE batch (the reference loops frame by frame at batch 1
    through PIL on the host, :30-34,44) — ToPILImage -> Resize(224) -> ToTensor becomes a GPU
    kernel that reproduces PIL's 8-bit two-pass bilinear resampler;
  * the trunk (frozen + eval when pretrained, :11-14) runs with BatchNorm folded into the
    convolutions: 7x7 stem = im2col + tcgen05 GEMM, 1x1 convolutions = tcgen05 GEMMs with
    bias + ReLU epilogues, 3x3 = halo-mode implicit GEMM, residual add + ReLU / pools / average
    pool = vectorised kernels. It is forward-only, as in the reference's pretrained configuration;
  * `linear` (the only trainable part) runs in fp32 with full autograd support, and the 3x16x16
    tiles are pasted into the 5x5 mosaic by a kernel (:36-38).

A trunk in train mode (batch-statistics BatchNorm at batch 1, trainable trunk) is not part of the
hot path (every reference driver constructs it with pretrained=True, rovr/rovr.py:31,
rovr/imitation_learning.py:39) and raises NotImplementedError.
"""
import torch
import torchvision.models as models
import torchvision.transforms as transforms

import ops
from _heads import LinearF32

BF = torch.bfloat16


class _MosaicPaste(torch.autograd.Function):
    """feature rows [b*S, ch*tile*tile] -> feature_map [b, ch, 5*tile, 5*tile] (zeros elsewhere),
    rovr/resnet_extractor.py:28-38 (ch = 3, tile = 16; video_processor.VideoProcessor: ch = 1, tile = 32)."""

    @staticmethod
    def forward(ctx, feat, b, S, ch=3, tile=16):
        feat = feat.contiguous()
        fmap = torch.zeros((b, ch, 5 * tile, 5 * tile), dtype=torch.float32, device=feat.device)
        ops.mosaic_paste(feat, fmap, slots_per_mosaic=S, tile=tile)
        ctx.S, ctx.tile = S, tile
        ctx.n = feat.shape
        return fmap

    @staticmethod
    def backward(ctx, g):
        gfeat = torch.empty(ctx.n, dtype=torch.float32, device=g.device)
        ops.mosaic_paste(gfeat, g.contiguous().float(), slots_per_mosaic=ctx.S, tile=ctx.tile, gather=True)
        return gfeat, None, None, None, None


class ResnetFeatureExtractor(torch.nn.Module):
    """Reference: rovr/resnet_extractor.py:5-67."""

    def __init__(self, pretrained=False):
        super().__init__()
        self.resnet = models.resnet50(pretrained=pretrained)
        self.linear = torch.nn.Linear(2048, 16 * 16 * 3)
        if pretrained:
            self.resnet.eval()
            for param in self.resnet.parameters():
                param.requires_grad = False
        self.resnet = torch.nn.Sequential(*(list(self.resnet.children())[:-1]))
        # kept for API parity (rovr/rovr.py:106 calls it per frame on the host); forward()/encode()
        # use the equivalent GPU kernel instead
        self.preprocessing = transforms.Compose([
            transforms.ToPILImage(),
            transforms.Resize((224, 224)),
            transforms.ToTensor(),
        ])
        self._folded = {}
        self._bn_train = False

    # -- folded / packed trunk weights --------------------------------------------------------------
    def _conv_bn(self, conv, bn, kind):
        key = id(conv)
        ver = tuple((t.data_ptr(), t._version) for t in (conv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var))
        hit = self._folded.get(key)
        if hit is not None and hit[0] == ver:
            return hit[1], hit[2]
        wf, bias = ops.fold_bn(conv.weight.detach().float().contiguous(), bn.weight.detach(), bn.bias.detach(),
                               bn.running_mean, bn.running_var, bn.eps)
        if kind == "c3":
            wk = ops.repack_conv3x3(wf, False)
        else:                                   # 1x1 convolution or the im2col'ed 7x7 stem: [Cout, K]
            wk = ops.repack_linear(wf.reshape(wf.shape[0], -1), False)
        self._folded[key] = (ver, wk, bias)
        return wk, bias

    def _conv_raw(self, conv, kind):
        """Train-mode trunk: the convolution's own weights (no BatchNorm folded in, no bias), packed."""
        key = ("raw", id(conv))
        ver = (conv.weight.data_ptr(), conv.weight._version)
        hit = self._folded.get(key)
        if hit is not None and hit[0] == ver:
            return hit[1], None
        wf = conv.weight.detach().float().contiguous()
        wk = ops.repack_conv3x3(wf, False) if kind == "c3" else ops.repack_linear(wf.reshape(wf.shape[0], -1), False)
        self._folded[key] = (ver, wk, None)
        return wk, None

    def _weights(self, conv, bn, kind):
        return self._conv_raw(conv, kind) if self._bn_train else self._conv_bn(conv, bn, kind)

    def _bn(self, z, bn, relu):
        """Train-mode trunk only: z is the fp32 convolution output (the statistics and the centring must see the
        unrounded values: |mean| is a multiple of the standard deviation behind a ReLU, so a bf16 z would cost that
        multiple in relative accuracy at every one of the 53 layers); returns the bf16 activation
        BatchNorm(z) [+ReLU] with per-frame batch statistics."""
        if not self._bn_train:
            return z
        if bn.momentum is None or not bn.track_running_stats:
            raise NotImplementedError("ResnetFeatureExtractor (B200): train-mode BatchNorm needs momentum and "
                                      "track_running_stats, as torchvision's resnet50 has them")
        y = torch.empty(z.shape, dtype=BF, device=z.device)
        ops.bn_train_fwd_frames(z, y, bn.weight.detach(), bn.bias.detach(), bn.eps, bn.momentum, bn.running_mean,
                                bn.running_var, bn.num_batches_tracked, relu=relu)
        return y

    def _check_trunk_mode(self):
        """Returns True for a train-mode trunk (the default constructor, rovr/resnet_extractor.py:6-8), False for the
        frozen eval-mode one (pretrained=True, :11-14 — what rovr.py:31 and imitation_learning.py:39 build)."""
        # decided from the BatchNorm layers themselves: `self.resnet` is a fresh nn.Sequential built AFTER the
        # reference's `.eval()` (rovr/resnet_extractor.py:11-16), so its own `training` flag is True even
        # though every child is in eval mode
        modes = {m.training for m in self.resnet.modules() if isinstance(m, torch.nn.modules.batchnorm._BatchNorm)}
        if len(modes) > 1:
            raise NotImplementedError("ResnetFeatureExtractor (B200): the trunk's BatchNorm layers must all be in the "
                                      "same mode")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.resnet.parameters()):
            raise NotImplementedError(
                "ResnetFeatureExtractor (B200): the trunk is forward-only; freeze it (requires_grad=False) as the "
                "reference does for pretrained=True, or call under torch.no_grad()")
        return True in modes

    def _c1(self, x, conv, bn, relu):
        if conv.stride[0] > 1:
            x = ops.subsample(x, conv.stride[0])
        B, H, W, C = x.shape
        wk, bias = self._weights(conv, bn, "c1")
        y = ops.gemm_bf16(x.reshape(-1, C), wk, bias, relu=relu and not self._bn_train,
                          out_dtype=torch.float32 if self._bn_train else BF)
        return self._bn(y.view(B, H, W, conv.out_channels), bn, relu)

    def _c3(self, x, conv, bn, relu):
        B, H, W, C = x.shape
        wk, bias = self._weights(conv, bn, "c3")
        s2 = conv.stride[0] == 2    # native stride 2: the TMA operand map samples every other input pixel
        shape = (B, (H - 1) // 2 + 1, (W - 1) // 2 + 1, conv.out_channels) if s2 else (B, H, W, conv.out_channels)
        if self._bn_train:
            z = torch.empty(shape, dtype=torch.float32, device=x.device)
            (ops.conv3x3_fprop_s2_f32out if s2 else ops.conv3x3_f32out)(x, wk, None, z, relu=False)
            return self._bn(z, bn, relu)
        y = torch.empty(shape, dtype=BF, device=x.device)
        return (ops.conv3x3_fprop_s2 if s2 else ops.conv3x3_fprop)(x, wk, bias, y, relu=relu)

    def _trunk(self, frames, bn_train=False):
        """frames: NCHW fp32 [n,3,224,224] already ToTensor-quantised -> pooled features [n,2048] fp32.
        bn_train: every BatchNorm uses (and records) per-frame batch statistics — the reference's trunk in
        training mode at its batch of one frame per call — instead of being folded into the convolution."""
        self._bn_train = bool(bn_train)
        r = self.resnet
        conv1, bn1, layers = r[0], r[1], (r[4], r[5], r[6], r[7])
        n = frames.shape[0]
        cols, Ho, Wo = ops.stem_im2col(frames, 160, quantise=False)
        wk, bias = self._weights(conv1, bn1, "stem")
        x = self._bn(ops.gemm_bf16(cols, wk, bias, relu=not self._bn_train,
                                   out_dtype=torch.float32 if self._bn_train else BF).view(n, Ho, Wo, 64), bn1, True)
        x = ops.maxpool_pad_fwd(x, 3, 2, 1)
        for layer in layers:
            for blk in layer:
                idt = x
                y = self._c1(x, blk.conv1, blk.bn1, True)
                y = self._c3(y, blk.conv2, blk.bn2, True)
                y = self._c1(y, blk.conv3, blk.bn3, False)
                if blk.downsample is not None:
                    idt = self._c1(x, blk.downsample[0], blk.downsample[1], False)
                x = ops.add_relu(y, idt)
        return ops.avgpool(x)

    def _preprocess_gpu(self, frames):
        """ToPILImage -> Resize((224,224)) -> ToTensor (rovr/resnet_extractor.py:18-23) for a batch."""
        return ops.resize_antialias(frames.float().contiguous(), 224, 224)

    def _encode_batch(self, frames):
        """frames [n,3,h,w] in [0,1] -> [n,768] features (grad flows into `linear` only)."""
        if not frames.is_cuda:
            raise RuntimeError("ResnetFeatureExtractor (B200) needs CUDA tensors: there is no CPU path")
        bn_train = self._check_trunk_mode()
        with torch.no_grad():
            pooled = self._trunk(self._preprocess_gpu(frames), bn_train)
        return LinearF32.apply(pooled, self.linear.weight, self.linear.bias)

    # -- reference API --------------------------------------------------------------------------------
    def forward(self, x):
        batch_size, seq_len, c, h, w = x.size()
        feats = self._encode_batch(x.reshape(batch_size * seq_len, c, h, w))
        return _MosaicPaste.apply(feats, batch_size, seq_len)

    def calculate_index(self, idx):
        return (idx // 5 * 16, idx % 5 * 16)

    def encode(self, x):
        return self._encode_batch(x.unsqueeze(0)).view(3, 16, 16)

    def insert_encoded_frame_batch(self, indices, full_frame_batch, encoded_frame_batch):
        feats = self._encode_batch(full_frame_batch)
        for b in range(full_frame_batch.size(0)):
            idx = self.calculate_index(indices[b])
            encoded_frame_batch[b, :, idx[0]:idx[0] + 16, idx[1]:idx[1] + 16] = feats[b].view(3, 16, 16)
        return encoded_frame_batch

    def extract_patch(self, indices, feature_map):
        local_device = feature_map.device
        patches = []
        for b, batch_indices in enumerate(indices):
            batch_patches = []
            for idx in batch_indices:
                i, j = self.calculate_index(idx)
                batch_patches.append(feature_map[b, :, i:i + 16, j:j + 16].unsqueeze(0))
            patches.append(torch.cat(batch_patches, 0))
        return torch.stack(patches).to(local_device)
