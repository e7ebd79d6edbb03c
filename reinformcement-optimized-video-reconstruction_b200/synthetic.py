"""Synthetic masked-video batches of the shape the reference's training scripts use.

The reference trains on RealVSR clips decoded by cv2 (rovr/video_ds.py), a dataset that is not
shipped; the benchmark and the examples use this generator instead. It mirrors
the dataset's structure: 25-frame clips, 256x256 RGB in [0,1], a zeroed 150 x 100 box per frame
whose position follows a raster scan (rovr/video_ds.py:19,62-87), model input = corrupted frame f
plus corrupted frames f-2, f-1 as context, target = clean frame f-1
(rovr/train_local_net_unet.py:44-52).
"""
import torch
import torch.nn.functional as F


def masked_frame_batch(B, H=256, W=256, seed=1234):
    """Returns CPU fp32 tensors: frame [B,3,H,W], context [B,2,3,H,W], target [B,3,H,W]."""
    gen = torch.Generator().manual_seed(seed)
    base = torch.rand((B, 3, 16, 16), generator=gen)
    big = F.interpolate(base, size=(H + 8, W + 8), mode="bilinear", align_corners=False)
    fidx = torch.randint(2, 25, (B,), generator=gen)
    bw, bh = max(1, (150 * W) // 256), max(1, (100 * H) // 256)

    def frame_at(b, n):
        dx, dy = int(n) % 8, (int(n) // 3) % 8
        clean = big[b, :, dy:dy + H, dx:dx + W]
        mask = torch.ones((1, H, W))
        x0 = ((int(n) % 8) * 32 * W) // 256
        y0 = ((int(n) // 8) * (256 // 3) * H) // 256
        mask[:, y0:min(H, y0 + bh), x0:min(W, x0 + bw)] = 0.0
        return clean, clean * mask

    frames, ctx, tgt = [], [], []
    for b in range(B):
        f = int(fidx[b])
        _, cf = frame_at(b, f)
        clean_m, cm = frame_at(b, f - 1)
        _, cn = frame_at(b, f - 2)
        frames.append(cf)
        ctx.append(torch.stack([cn, cm], 0))
        tgt.append(clean_m)
    return (torch.stack(frames).contiguous(), torch.stack(ctx).contiguous(),
            torch.stack(tgt).contiguous())
