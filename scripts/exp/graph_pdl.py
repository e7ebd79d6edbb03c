"""Does the captured step keep its programmatic-dependent-launch edges? Dumps the CUDA graph as DOT and counts edge kinds;
times graph replay vs eager launches with and without PDL (ROVR_PDL is read at library load: run once per setting)."""
import os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "reinformcement-optimized-video-reconstruction_b200"))
import torch
from local_net import LocalNetworkUNetNorm, _forward_impl, _backward_impl, _make_buckets
from synthetic import masked_frame_batch

dev = torch.device("cuda:0")
torch.manual_seed(0)
net = LocalNetworkUNetNorm().to(dev)
x, c, t = [v.to(dev) for v in masked_frame_batch(24, 256, 256, seed=1)]
named = dict(net.named_parameters())
Pd = {n: named[n].detach() for n in net._live_names}
flats, G = _make_buckets(Pd, dev)
one = torch.ones((), device=dev)


def run():
    y, loss, acts = _forward_impl(net, x, c, t, Pd, True)
    _backward_impl(net, acts, Pd, t, None, one, buckets=(flats, G))
    return loss


def timed(fn, n=30):
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


s = torch.cuda.Stream()
with torch.cuda.stream(s):
    run(); run()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
g.enable_debug_mode()
with torch.cuda.graph(g):
    run()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
dot = os.path.join(ROOT, "gpurun_out", f"graph_pdl{os.environ.get('ROVR_PDL', '1')}.dot")
print("ROVR_PDL =", os.environ.get("ROVR_PDL", "1"))
try:
    g.debug_dump(dot)
    txt = open(dot).read()
    edges = re.findall(r"->.*", txt)
    print("nodes", len(re.findall(r"label=", txt)), "| edges", len(edges),
          "| edges mentioning programmatic/port:", sum(1 for e in edges if re.search(r"rogrammatic|port|PROGRAMMATIC", e)))
    print([e for e in edges][:6])
except Exception as exc:
    print("no dot dump:", exc)
print("eager  ms/step", timed(run))
print("replay ms/step", timed(g.replay))
