import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "reinformcement-optimized-video-reconstruction_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch, torch.nn.functional as F
import rovr_oracle as O, ops
from _blocks import TrunkOps, PackedWeights, BF
dev = torch.device("cuda:0")
sd = O.pn2_state_dict(0, False)
g = torch.Generator().manual_seed(21)
b = 20
enc = torch.rand((b, 1, 160, 160), generator=g)
def rel(a, r): a=a.float().cpu(); r=r.float().cpu(); return ((a-r).norm()/(r.norm()+1e-20)).item()
def nchw(t): return t.float().permute(0,3,1,2)
# ---- emulation with intermediates
rt = lambda t, gr=True: O._RoundTrip.apply(t, gr)
wq = lambda p: p.to(torch.bfloat16).float()
E = {}
def keep(n, t): t.retain_grad(); E[n] = t; return t
x = rt(enc, False)
def cbr(i, t, tag):
    c, bn = f"video_conv.{i}", f"video_conv.{i+1}"
    raw = keep("raw"+tag, rt(F.conv2d(t, wq(sd[c+".weight"]), sd[c+".bias"], padding=1)))
    return keep("a"+tag, rt(F.relu(O._bn_train(sd, bn, raw))))
x.requires_grad_(True)
a0 = cbr(0, x, "0"); p0 = keep("p0", F.max_pool2d(a0, 8, 8))
a4 = cbr(4, p0, "4"); p4 = keep("p4", F.max_pool2d(a4, 4, 4))
a8 = cbr(8, p4, "8"); a12 = cbr(12, a8, "12")
q = keep("q", F.max_pool2d(a12, 2, (2, 1))); r = keep("r", F.max_pool2d(q, 2, (2, 2)))
gout = torch.randn(r.shape, generator=g)
(r * gout).sum().backward()
# ---- CUDA
P = {k: v.to(dev) for k, v in sd.items() if v.is_floating_point() and "running" not in k}
Bf = {k: v.to(dev).clone() for k, v in sd.items() if "running" in k or "num_batches" in k}
T = TrunkOps(P, Bf, PackedWeights(), True)
VC = [("video_conv.0", "video_conv.1"), ("video_conv.4", "video_conv.5"), ("video_conv.8", "video_conv.9"), ("video_conv.12", "video_conv.13")]
def buf(h, w, c): return torch.empty((b, h, w, c), dtype=BF, device=dev)
a = {}
a["in16"] = ops.pack_nchw([enc.to(dev)], 16)
a["a0"] = buf(160,160,64); a["r0"] = T.cbr3_fwd(*VC[0], a["in16"], a["a0"]); a["p0"] = ops.maxpool_fwd(a["a0"], buf(20,20,64), 8)
a["a4"] = buf(20,20,128); a["r4"] = T.cbr3_fwd(*VC[1], a["p0"], a["a4"]); a["p4"] = ops.maxpool_fwd(a["a4"], buf(5,5,128), 4)
a["a8"] = buf(5,5,256); a["r8"] = T.cbr3_fwd(*VC[2], a["p4"], a["a8"])
a["a12"] = buf(5,5,512); a["r12"] = T.cbr3_fwd(*VC[3], a["a8"], a["a12"])
a["q"] = ops.maxpool_fwd(a["a12"], buf(2,4,512), 2, (2,1)); a["r"] = ops.maxpool_fwd(a["q"], buf(1,2,512), 2, (2,2))
for n, e in [("a0","a0"),("p0","p0"),("a4","a4"),("p4","p4"),("a8","a8"),("a12","a12"),("q","q"),("r","r")]:
    print("fwd", n, rel(nchw(a[n]), E[e]), "exact-equal frac", (nchw(a[n]).cpu()==E[e].detach()).float().mean().item())
print("raw12", rel(nchw(a["r12"][0]), E["raw12"]))
like = torch.empty_like
gr = gout.permute(0,2,3,1).contiguous().to(dev).to(BF)
print("gr in", rel(nchw(gr), gout))
gq = ops.maxpool_bwd(a["q"], gr, like(a["q"]), 2, (2,2), relu_mask=False); print("gq", rel(nchw(gq), E["q"].grad))
ga12 = ops.maxpool_bwd(a["a12"], gq, like(a["a12"]), 2, (2,1), relu_mask=False); print("ga12", rel(nchw(ga12), E["a12"].grad))
raw, mean, rstd = a["r12"]
draw = T._bn_bwd("video_conv.13", ga12, a["a12"], raw, mean, rstd, 512); print("draw12", rel(nchw(draw), E["raw12"].grad))
ga8 = like(a["a8"]); T.cbr3_bwd(*VC[3], a["r12"], a["a8"], a["a12"], ga12, ga8); print("ga8", rel(nchw(ga8), E["a8"].grad))
gp4 = like(a["p4"]); T.cbr3_bwd(*VC[2], a["r8"], a["p4"], a["a8"], ga8, gp4); print("gp4", rel(nchw(gp4), E["p4"].grad))
ga4 = ops.maxpool_bwd(a["a4"], gp4, like(a["a4"]), 4, relu_mask=False); print("ga4", rel(nchw(ga4), E["a4"].grad))
gp0 = like(a["p0"]); T.cbr3_bwd(*VC[1], a["r4"], a["p0"], a["a4"], ga4, gp0); print("gp0", rel(nchw(gp0), E["p0"].grad))
ga0 = ops.maxpool_bwd(a["a0"], gp0, like(a["a0"]), 8, relu_mask=False); print("ga0", rel(nchw(ga0), E["a0"].grad))
