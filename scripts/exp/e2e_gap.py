"""Where do the ~0.2 ms between the resident graph replay and the end-to-end graphed loop go? (experiment)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "reinformcement-optimized-video-reconstruction_b200"))
import torch
import ops
from feeder import DeviceFeeder, ScalarReadback
from local_net import GraphedTrainingStep, LocalNetworkUNetNorm
from synthetic import masked_frame_batch

dev = torch.device("cuda:0")
torch.manual_seed(0)
net = LocalNetworkUNetNorm().to(dev)
host = [(v * 255).round().to(torch.uint8).pin_memory() for v in masked_frame_batch(24, 256, 256, seed=1)]
x, c, t = [v.to(dev).float().div_(255) for v in host]
u8 = [v.to(dev) for v in host]
step = GraphedTrainingStep(net, x, c, t)
N = 30


def timed(fn, n=N):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


print("replay only                       ", timed(lambda: step()))
print("replay + 3 x u8->f32 (resident u8)", timed(lambda: step(*u8)))
print("replay + 3 x fp32 copy_           ", timed(lambda: step(x, c, t)))


def loop(to_float, readback, n=N):
    rb = ScalarReadback(dev, lag=1)
    for xd, cd, td in DeviceFeeder((tuple(host) for _ in range(n)), dev, to_float=to_float):
        loss = step(xd, cd, td)
        if readback:
            rb.exchange(loss)
    rb.drain()


for tf in (False, True):
    for rbk in (False, True):
        loop(tf, rbk, 3)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        loop(tf, rbk)
        e1.record()
        torch.cuda.synchronize()
        print(f"feeder to_float={tf} readback={rbk}: ", e0.elapsed_time(e1) / N)
# CPU time of one call
t0 = time.perf_counter()
for _ in range(N):
    step(*u8)
cpu = (time.perf_counter() - t0) / N
torch.cuda.synchronize()
print("CPU time per step(u8) call (ms)", cpu * 1e3)
