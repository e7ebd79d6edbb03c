"""Per-call CUDA-event timing of one EncoderBlock forward + backward (E=3072, S=256, B=8, 8 heads)."""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "reinformcement-optimized-video-reconstruction_b200"))
import torch, ops
from common_layers import EncoderBlock
dev = torch.device("cuda:0")
E, S, B = 3072, 256, int(os.environ.get("B", "8"))
blk = EncoderBlock(E, 8, 0.0).to(dev).eval()
x = torch.randn(B, S, E, device=dev, requires_grad=True)
def step():
    blk.zero_grad(set_to_none=True); x.grad = None
    (blk(x) ** 2).sum().backward()
for _ in range(3): step()
torch.cuda.synchronize()
prof = []
ops.set_profile(prof)
for _ in range(5): step()
ops.set_profile(None)
torch.cuda.synchronize()
agg = collections.OrderedDict()
for name, a, b in prof:
    d = agg.setdefault(name, [0, 0.0]); d[0] += 1; d[1] += a.elapsed_time(b)
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:28s} calls/step {v[0]/5:6.1f}  ms/step {v[1]/5:7.3f}  {100*v[1]/tot:5.1f}%")
print("total ms/step (sum of calls)", tot / 5)
