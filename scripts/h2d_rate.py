"""Pinned host -> device copy rate at the bench's per-step input size (75.5 MB)."""
import torch
dev = torch.device("cuda:0")
h = torch.empty(75497472 // 4, dtype=torch.float32).pin_memory()
d = torch.empty_like(h, device=dev)
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        d.copy_(h, non_blocking=True)
    e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"H2D 75.5 MB: {ms:.3f} ms  {75.497472 / ms:.1f} GB/s")
