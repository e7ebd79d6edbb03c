"""CPU experiment (not product code): which storage / operand precisions of the policy trunks give
gradients within 2e-2 of the fp32 reference?  Variants emulated in PyTorch fp32 arithmetic:
  fwd operand quantiser qf in {bf16, split2 (hi+lo bf16 = 16-bit mantissa), fp32}
  bwd operand quantiser qb in {bf16, split2}
  activation storage for BN/pool: fp32 always here (that is the proposed fix)
"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "oracle"))
import torch, torch.nn.functional as F
import rovr_oracle as O

def bf(t): return t.to(torch.bfloat16).to(torch.float32)
def split2(t):
    hi = bf(t); return hi + bf(t - hi)
Q = {"bf16": bf, "split2": split2, "fp32": lambda t: t}

class QConv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, pad, transposed, qf, qb):
        ctx.save_for_backward(x, w); ctx.cfg = (pad, transposed, qb)
        xq, wq = Q[qf](x), Q[qf](w)
        return F.conv_transpose2d(xq, wq, b, stride=2) if transposed else F.conv2d(xq, wq, b, padding=pad)
    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors; pad, transposed, qb = ctx.cfg
        q = Q[qb]; xq, wq, gq = q(x), q(w), q(g)
        with torch.enable_grad():
            xx = xq.detach().requires_grad_(True); ww = wq.detach().requires_grad_(True)
            y = F.conv_transpose2d(xx, ww, None, stride=2) if transposed else F.conv2d(xx, ww, None, padding=pad)
            gx, gw = torch.autograd.grad(y, (xx, ww), gq)
        return gx, gw, g.sum((0, 2, 3)), None, None, None, None

def pn1_unet(sd, x, qf, qb):
    def cbr(conv, bn, t, pad=1):
        return F.relu(O._bn_train(sd, bn, QConv.apply(t, sd[conv + ".weight"], sd[conv + ".bias"], pad, False, qf, qb)))
    def ubr(up, bn, t):
        return F.relu(O._bn_train(sd, bn, QConv.apply(t, sd[up + ".weight"], sd[up + ".bias"], 0, True, qf, qb)))
    e1 = cbr("conv1", "bn1", x); e2 = cbr("conv2", "bn2", F.max_pool2d(e1, 2)); e3 = cbr("conv3", "bn3", F.max_pool2d(e2, 2))
    e4 = cbr("conv4", "bn4", F.max_pool2d(e3, 2))
    d = cbr("conv5", "bn5", torch.cat([ubr("upconv1", "bn_up1", e4), e3], 1))
    d = cbr("conv6", "bn6", torch.cat([ubr("upconv2", "bn_up2", d), e2], 1))
    d = cbr("conv7", "bn7", torch.cat([ubr("upconv3", "bn_up3", d), e1], 1))
    d = cbr("conv8", "bn8", d, pad=0); d = cbr("conv9", "bn9", F.max_pool2d(d, 2), pad=0)
    return F.max_pool2d(d, 2)

def pn2_vc(sd, image, qf, qb):
    def cbr(i, t):
        c, b = f"video_conv.{i}", f"video_conv.{i + 1}"
        return F.relu(O._bn_train(sd, b, QConv.apply(t, sd[c + ".weight"], sd[c + ".bias"], 1, False, qf, qb)))
    t = F.max_pool2d(cbr(0, image), 8, 8); t = F.max_pool2d(cbr(4, t), 4, 4); t = cbr(8, t); t = cbr(12, t)
    t = F.max_pool2d(t, 2, (2, 1)); t = F.max_pool2d(t, 2, (2, 2))
    return t.flatten(1)

def leaf(sd): return {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v) for k, v in sd.items()}
def rel(a, b): return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()

def run_pn1(b, qf, qb, critic):
    sd = O.pn1_state_dict(0, critic)
    g = torch.Generator().manual_seed(11)
    image = torch.rand((b, 3, 80, 80), generator=g); context = torch.rand((b, 3, 80, 80), generator=g)
    L = leaf(sd)
    feat = pn1_unet(L, torch.cat([image, context], 1), qf, qb).flatten(1)
    feat = (feat - feat.mean(1, keepdim=True)) / feat.std(1, keepdim=True)
    out = F.linear(feat, L["fc_final.weight"], L["fc_final.bias"])
    (out ** 2).sum().backward()
    return out.detach(), {k: v.grad for k, v in L.items() if v.requires_grad and v.grad is not None}

def run_pn2(b, qf, qb, critic):
    sd = O.pn2_state_dict(0, critic)
    g = torch.Generator().manual_seed(21)
    enc = torch.rand((b, 1, 160, 160), generator=g); feat = torch.randn((b, 1024), generator=g)
    L = leaf(sd)
    st = torch.cat([pn2_vc(L, enc, qf, qb), feat], 1)
    if critic:
        st = (st - st.mean(0, keepdim=True)) / (st.std(0, keepdim=True) + 0.001)
    out = O.pn2_final_fc(L, st)
    (out ** 2).sum().backward()
    return out.detach(), {k: v.grad for k, v in L.items() if v.requires_grad and v.grad is not None}

if __name__ == "__main__":
    torch.set_num_threads(16)
    for name, fn, b in (("pn1", run_pn1, 25), ("pn2", run_pn2, 20)):
        for critic in (False, True):
            o0, g0 = fn(b, "fp32", "fp32", critic)
            for qf, qb in (("bf16", "bf16"), ("split2", "bf16"), ("split2", "split2")):
                o, g = fn(b, qf, qb, critic)
                worst = max(((rel(g[k], g0[k]), k) for k in g0 if g0[k].norm() > 1e-6 * max(v.norm() for v in g0.values())))
                print(f"{name} critic={critic} fwd={qf} bwd={qb}: out rel {rel(o, o0):.2e} worst grad {worst[0]:.2e} ({worst[1]})", flush=True)
