import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "reinformcement-optimized-video-reconstruction_b200"))
import torch, ops
dev = torch.device("cuda:0")
BF = torch.bfloat16
def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
B = 24
for name, H, Cin, Cout in [("conv7", 256, 128, 64), ("conv6", 128, 256, 128), ("conv5", 64, 512, 256)]:
    dy = torch.randn(B, H, H, Cout, device=dev).to(BF)
    w = torch.randn(Cout, Cin, 3, 3, device=dev) * 0.05
    wd = ops.repack_conv3x3(w, True)
    mask = torch.randn(B, H, H, Cin, device=dev).clamp_min(0).to(BF)
    dx = torch.empty(B, H, H, Cin, device=dev, dtype=BF)
    cs = torch.empty(Cin // 2, device=dev)
    flops = 2 * B * H * H * 9 * Cin * Cout
    for tag, kw in [("plain", {}), ("mask", {"mask": mask}), ("colsum", {"colsum": cs}), ("both", {"mask": mask, "colsum": cs})]:
        us = timeit(lambda: ops.conv3x3_dgrad(dy, wd, dx, **kw))
        print(f"{name} dgrad {tag:7s} {us:8.1f} us  {flops / us / 1e6:7.1f} TFLOP/s")
    # fprop of the same layer for comparison
    x = torch.randn(B, H, H, Cin, device=dev).to(BF)
    wk = ops.repack_conv3x3(w, False)
    y = torch.empty(B, H, H, Cout, device=dev, dtype=BF)
    bias = torch.zeros(Cout, device=dev)
    us = timeit(lambda: ops.conv3x3_fprop(x, wk, bias, y))
    print(f"{name} fprop         {us:8.1f} us  {flops / us / 1e6:7.1f} TFLOP/s")
