"""`video_processor.VideoProcessor` — a RESTATEMENT.

rovr/rovr.py:16,61,107,200 and rovr/imitation_learning.py:19,58,78 import and call
`VideoProcessor`, but the file is NOT part of the reference repository: there is nothing to be
faithful to except its call sites, so parity is UNPINNED and un-oracled by the reference. What the
call sites fix:

    encoded_frames, flattened_frames = VideoProcessor()(stacked_frames)     # stacked [1, S, 3, 224, 224]
    encoded_frames = vp.insert_encoded_frame_batch(torch.tensor(j).view(-1, 1), image, encoded_frames)
    pn2(encoded_frames [b,1,160,160], flattened_frames-row [b,1,1024], target)   # 1024 + 1024 = final_fc's 2048

so `encoded_frames` is a [b, 1, 160, 160] mosaic that PolicyNetwork2UNet.video_conv reduces to 1024
features and `flattened_frames` carries one 1024-vector per frame. This module defines it as the
32-pixel-tile sibling of `ResnetFeatureExtractor` (rovr/resnet_extractor.py:25-55, which builds the
[b, 3, 80, 80] mosaic of 3x16x16 tiles the same way): frozen eval-mode ResNet-50 trunk -> 2048-d
pooled feature -> Linear(2048, 1024) -> `flattened_frames[b, S, 1024]`; each vector viewed as a
1x32x32 tile pasted at (idx // 5 * 32, idx % 5 * 32). All frames of a call are encoded as one batch
on the B200 kernels (the trunk, the fp32 projection and the mosaic scatter of resnet_extractor.py).
"""
import torch

import ops
from _heads import LinearF32
from resnet_extractor import ResnetFeatureExtractor, _MosaicPaste


class VideoProcessor(ResnetFeatureExtractor):
    TILE, CH = 32, 1

    def __init__(self, pretrained=False):
        super().__init__(pretrained=pretrained)
        self.linear = torch.nn.Linear(2048, self.CH * self.TILE * self.TILE)        # 1024-d frame vectors
        if not pretrained:
            # the drivers construct it without arguments and never train the trunk: frozen + eval, like the
            # pretrained extractor (rovr/resnet_extractor.py:11-14)
            self.resnet.eval()
            for p in self.resnet.parameters():
                p.requires_grad = False

    def train(self, mode=True):
        super().train(mode)
        self.resnet.eval()                      # the trunk stays in eval mode (frozen BatchNorm statistics)
        return self

    def forward(self, x):
        """x [b, S, 3, h, w] (S <= 25) -> (encoded_frames [b, 1, 160, 160], flattened_frames [b, S, 1024])."""
        b, S, c, h, w = x.size()
        feats = self._encode_batch(x.reshape(b * S, c, h, w))
        side = 5 * self.TILE
        return _MosaicPaste.apply(feats, b, S, self.CH, self.TILE), feats.view(b, S, -1)

    def calculate_index(self, idx):
        return (idx // 5 * self.TILE, idx % 5 * self.TILE)

    def encode(self, x):
        return self._encode_batch(x.unsqueeze(0)).view(self.CH, self.TILE, self.TILE)

    def insert_encoded_frame_batch(self, indices, full_frame_batch, encoded_frame_batch):
        """rovr/rovr.py:200: re-encode the reconstructed frame(s) and overwrite their tiles in place."""
        feats = self._encode_batch(full_frame_batch)
        T = self.TILE
        for b in range(full_frame_batch.size(0)):
            i, j = self.calculate_index(int(indices[b]))
            encoded_frame_batch[b, :, i:i + T, j:j + T] = feats[b].view(self.CH, T, T)
        return encoded_frame_batch
