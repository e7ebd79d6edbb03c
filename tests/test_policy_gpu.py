"""GPU parity tests of the policy-network path: BatchNorm / linear / head kernels against plain
PyTorch fp32, and the PolicyNetwork1UNet / PolicyNetwork2UNet / ActionLSTM drop-ins against the
oracle (oracle/rovr_oracle.py) and the golden fixtures produced by the unmodified reference.

Tolerances (north_star): fp32 head kernels 1e-4 relative; the policy trunks run on the
emulated-fp32 tensor-core path (split-bf16 products, csrc/fp32x.cuh) and are held to **2e-2 on every
parameter gradient and on both critic values against the pure-fp32 oracle** (measured ~1e-4; PN1
at the config's b = 25); selected frame indices bit-exact against the reference's golden indices.

The plain bf16-operand trunks stay available (`trunk_precision = "bf16"`, half the memory): their
gradients behind the max-pools deviate from fp32 by 0.2-0.6 L2-rel because bf16 products flip the
arg-max of a few percent of the pooling windows — a property of bf16 arithmetic that the oracle's
own bf16-storage emulation reproduces without any CUDA code; test_policy_bf16_mode_bounds keeps
that mode honest against the emulation with its documented bounds.
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import rovr_oracle as O

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def _dev():
    import _native
    _native.require_device()
    return torch.device("cuda:0")


def _rel(got, ref):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    return ((got - ref).norm() / (ref.norm() + 1e-20)).item()


def _close(name, got, ref, tol):
    r = _rel(got, ref)
    print(f"{name}: l2-rel {r:.3e}")
    assert r < tol and not torch.isnan(got).any(), f"{name}: l2-rel {r:.3e} >= {tol}"


# ------------------------------------------------------------------------------------------------
# kernels
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,H,W,C,cv", [(5, 80, 80, 32, 32), (3, 20, 20, 128, 128), (7, 5, 5, 512, 512),
                                        (2, 40, 40, 16, 3), (4, 24, 16, 16, 1), (1, 80, 80, 64, 64)])
def test_bn_train_fwd_bwd(B, H, W, C, cv):
    import ops
    dev = _dev()
    g = torch.Generator().manual_seed(C + H)
    x = (torch.randn((B, H, W, C), generator=g) * 1.5 + 0.3).to(dev).to(BF)
    x[..., cv:] = 0
    gamma = (1 + 0.2 * torch.randn(cv, generator=g)).to(dev)
    beta = (0.2 * torch.randn(cv, generator=g)).to(dev)
    rm, rv = torch.zeros(cv, device=dev), torch.ones(cv, device=dev)
    nbt = torch.zeros((), dtype=torch.long, device=dev)
    ybuf = torch.full((B, H, W, C + 8), 3.0, dtype=BF, device=dev)
    y = ybuf[..., 8:]
    mean, rstd = ops.bn_train_fwd(x, y, gamma, beta, cv, 1e-5, 0.1, rm, rv, nbt, relu=True)
    xr = x.float()[..., :cv].permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    rm_ref, rv_ref = torch.zeros(cv, device=dev), torch.ones(cv, device=dev)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    yr = F.relu(F.batch_norm(xr, rm_ref, rv_ref, gr, br, True, 0.1, 1e-5))
    _close("bn.y", y[..., :cv].float().permute(0, 3, 1, 2), yr, 6e-3)
    assert (y[..., cv:] == 0).all() and (ybuf[..., :8] == 3.0).all()
    _close("bn.running_mean", rm, rm_ref, 1e-4)
    _close("bn.running_var", rv, rv_ref, 1e-4)
    assert int(nbt) == 1
    dy = torch.randn((B, H, W, C), generator=g).to(dev).to(BF)
    dx = torch.empty_like(x)
    dgam, dbet = torch.empty(cv, device=dev), torch.empty(cv, device=dev)
    ops.bn_train_bwd(dy, y, x, dx, gamma, mean, rstd, cv, dgam, dbet, relu=True)
    # reference backward with the SAME relu mask (the stored bf16 activation)
    mask = (y[..., :cv].float() > 0).permute(0, 3, 1, 2)
    yr2 = F.batch_norm(xr, None, None, gr, br, True, 0.0, 1e-5)
    (yr2 * mask * dy.float()[..., :cv].permute(0, 3, 1, 2)).sum().backward()
    _close("bn.dx", dx[..., :cv].float().permute(0, 3, 1, 2), xr.grad, 1e-2)
    _close("bn.dgamma", dgam, gr.grad, 2e-3)
    _close("bn.dbeta", dbet, br.grad, 2e-3)
    assert (dx[..., cv:] == 0).all()


@pytest.mark.parametrize("M,N,K", [(20, 1024, 2048), (25, 25, 400), (1, 19200, 1024), (2, 256, 2307), (20, 20, 64),
                                   (25, 768, 2048)])
def test_linear_f32(M, N, K):
    import ops
    dev = _dev()
    g = torch.Generator().manual_seed(M + N + K)
    x = torch.randn((M, K), generator=g).to(dev)
    w = (torch.randn((N, K), generator=g) / K ** 0.5).to(dev)
    b = torch.randn(N, generator=g).to(dev)
    y = ops.linear_f32_fwd(x, w, b)
    ref = (x.double() @ w.double().t() + b.double()).float()
    _close("linear.fwd", y, ref, 1e-5)
    dy = torch.randn((M, N), generator=g).to(dev)
    dx = ops.linear_f32_dgrad(dy, w)
    _close("linear.dgrad", dx, (dy.double() @ w.double()).float(), 1e-5)
    dw, db = torch.empty_like(w), torch.empty_like(b)
    ops.linear_f32_wgrad(dy, x, dw, db)
    _close("linear.wgrad", dw, (dy.double().t() @ x.double()).float(), 1e-5)
    _close("linear.bgrad", db, dy.sum(0), 1e-5)


@pytest.mark.parametrize("M,N,K,nk,kk", [(12800, 16, 32, 3, 32), (3000, 16, 16, 1, 3), (512, 384, 128, 384, 128),
                                         (1000, 128, 512, 128, 512)])
def test_gemm_wgrad(M, N, K, nk, kk):
    import ops
    dev = _dev()
    g = torch.Generator().manual_seed(M + N)
    dy = torch.randn((M, N), generator=g).to(dev).to(BF)
    x = torch.randn((M, K), generator=g).to(dev).to(BF)
    dw = torch.empty((nk, kk), device=dev)
    ops.gemm_wgrad(dy, x, dw)
    ref = (dy.float().t() @ x.float())[:nk, :kk]
    _close("gemm_wgrad", dw, ref, 2e-3)


@pytest.mark.parametrize("R,C,dim,eps", [(25, 400, 1, 0.0), (20, 2048, 0, 0.001), (1, 400, 1, 0.0)])
def test_standardize(R, C, dim, eps):
    import ops
    dev = _dev()
    g = torch.Generator().manual_seed(R + C)
    x = (torch.randn((R, C), generator=g) * 2 + 0.5).to(dev)
    if dim == 0 and R < 2:
        pytest.skip("needs two rows")
    y, sig = ops.standardize_fwd(x, dim, eps)
    xr = x.clone().requires_grad_(True)
    yr = (xr - xr.mean(dim=dim, keepdim=True)) / (xr.std(dim=dim, keepdim=True) + eps)
    _close("standardize.y", y, yr, 1e-5)
    gy = torch.randn((R, C), generator=g).to(dev)
    yr.backward(gy)
    dx = ops.standardize_bwd(gy, y, sig, dim, eps)
    _close("standardize.dx", dx, xr.grad, 1e-4)


@pytest.mark.parametrize("b,n,tk", [(20, 20, 1), (1, 20, 1), (1, 25, 0), (25, 25, 0)])
def test_head_mask_std(b, n, tk):
    import ops
    dev = _dev()
    g = torch.Generator().manual_seed(b * 100 + n)
    l0 = torch.randn((b, n), generator=g).to(dev)
    target = torch.randint(0, n, (b, tk), generator=g).to(dev) if tk else None
    l = l0.clone()
    out, sig = ops.head_mask_std_fwd(l, target, True)
    lr = l0.clone().requires_grad_(True)
    lm = lr.scatter(1, target, 0.0) if tk else lr
    ref = (lm - lm.mean(dim=1)) / (lm.std(dim=(1,), keepdim=True) + 0.1)   # the reference's expression
    _close("mask_std.out", out, ref, 1e-5)
    gy = torch.randn((b, n), generator=g).to(dev)
    ref.backward(gy)
    dl = ops.head_mask_std_bwd(gy, l, out, sig, target, True)
    _close("mask_std.dl", dl, lr.grad, 1e-4)


def test_head_gumbel_modes():
    import ops
    dev = _dev()
    g = torch.Generator().manual_seed(5)
    for b, n, tau in [(20, 20, 0.7), (1, 20, 0.7), (5, 25, 0.5), (300, 25, 0.5)]:
        logits = torch.randn((b, n), generator=g).to(dev)
        expo = torch.empty((b, n)).exponential_(generator=g).to(dev)
        ref_p = O.gumbel_softmax_with_noise(logits, tau, expo)
        probs, _, _ = ops.head_gumbel_fwd(logits, expo, tau, 0)
        _close("gumbel.probs", probs, ref_p, 1e-5)
        _, idx1, val1 = ops.head_gumbel_fwd(logits, expo, tau, 1)
        mx = ref_p.max(dim=1)
        assert torch.equal(idx1, mx.indices), "argmax index differs"
        _close("gumbel.max.logp", val1, mx.values.log(), 1e-5)
        _, idx2, val2 = ops.head_gumbel_fwd(logits, expo, tau, 2)
        top = torch.topk(ref_p, 2, dim=1)
        assert torch.equal(idx2, top.indices), "top-2 indices differ"
        _close("gumbel.top2.logp", val2, top.values.log().sum(1) / 2 + 0.69314, 1e-5)
        # differentiable modes
        act1 = torch.randint(0, n, (b,), generator=g).to(dev)
        lr = logits.clone().requires_grad_(True)
        ref3 = O.gumbel_softmax_with_noise(lr, tau, expo).gather(1, act1[:, None]).log().squeeze(1)
        p3, _, v3 = ops.head_gumbel_fwd(logits, expo, tau, 3, act1)
        _close("gumbel.lp1", v3, ref3, 1e-5)
        gv = torch.randn(b, generator=g).to(dev)
        ref3.backward(gv)
        _close("gumbel.lp1.dl", ops.head_gumbel_bwd(p3, gv, tau, 3, act1), lr.grad, 1e-4)
        act2 = torch.randint(0, n, (b, 2), generator=g).to(dev)
        lr = logits.clone().requires_grad_(True)
        pr = O.gumbel_softmax_with_noise(lr, tau, expo)
        pair = (pr[:, :, None] * pr[:, None, :]).flatten(1)
        ref4 = pair.gather(1, (act2[:, 0] * n + act2[:, 1])[:, None]).log().sum(1) / 2 + 0.69314
        p4, _, v4 = ops.head_gumbel_fwd(logits, expo, tau, 4, act2)
        _close("gumbel.lp2", v4, ref4, 1e-5)
        ref4.backward(gv)
        _close("gumbel.lp2.dl", ops.head_gumbel_bwd(p4, gv, tau, 4, act2), lr.grad, 1e-4)


def test_lstm_pointwise():
    import ops
    dev = _dev()
    g = torch.Generator().manual_seed(9)
    B, Hd = 3, 64
    gates = torch.randn((B, 4 * Hd), generator=g).to(dev)
    c0 = torch.randn((B, Hd), generator=g).to(dev)
    h, c, act = ops.lstm_pointwise_fwd(gates, c0)
    gr, cr = gates.clone().requires_grad_(True), c0.clone().requires_grad_(True)
    i, f, gg, o = gr.chunk(4, 1)
    c_ref = torch.sigmoid(f) * cr + torch.sigmoid(i) * torch.tanh(gg)
    h_ref = torch.sigmoid(o) * torch.tanh(c_ref)
    _close("lstm.h", h, h_ref, 1e-5)
    _close("lstm.c", c, c_ref, 1e-5)
    dh, dc = torch.randn((B, Hd), generator=g).to(dev), torch.randn((B, Hd), generator=g).to(dev)
    torch.autograd.backward([h_ref, c_ref], [dh, dc])
    dg, dcp = ops.lstm_pointwise_bwd(act, c0, c, dh, dc)
    _close("lstm.dgates", dg, gr.grad, 1e-4)
    _close("lstm.dc_prev", dcp, cr.grad, 1e-4)


# ------------------------------------------------------------------------------------------------
# modules
# ------------------------------------------------------------------------------------------------
def _golden(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def _pn1_inputs(b, seed):
    g = torch.Generator().manual_seed(seed)
    image = torch.rand((b, 3, 80, 80), generator=g)
    context = torch.rand((b, 3, 80, 80), generator=g)
    action = torch.randint(0, 25, (b,), generator=g)
    return image, context, action


def _pn2_inputs(b, seed):
    g = torch.Generator().manual_seed(seed)
    enc = torch.rand((b, 1, 160, 160), generator=g)
    feat = torch.randn((b, 1, 1024), generator=g)
    target = torch.randint(0, 20, (b, 1, 1), generator=g)
    a0 = torch.randint(0, 20, (b,), generator=g)
    a1 = (a0 + 1 + torch.randint(0, 19, (b,), generator=g)) % 20
    return enc, feat, target, torch.stack([a0, a1], 1)


def _leaf(sd):
    return {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v)
            for k, v in sd.items()}


def _fixed_noise(monkeypatch, module, expo):
    """Make the module's gumbel draw return `expo` (the CPU draw the golden fixtures were made with)."""
    monkeypatch.setattr(module, "exponential_like", lambda logits: expo.to(logits.device).clone())


def _check_param_grads(tag, net, ref_grads, tol, skip_bias_before_bn=(), floor=1e-4):
    """Every gradient within l2-rel `tol` of the reference. `floor` (x the largest gradient norm of the
    network) is the absolute slack for gradients that are ~0 by cancellation (parameters whose effect a
    later standardisation removes): 1e-4 for the emulated-fp32 trunks, 2e-2 for the bf16 mode."""
    named = dict(net.named_parameters())
    worst = 0.0
    bad = []
    gmax = max(float(v.float().norm()) for v in ref_grads.values() if v is not None)
    for name, gr in ref_grads.items():
        if gr is None:
            assert named[name].grad is None, f"{name} should have no gradient"
            continue
        g = named[name].grad
        assert g is not None, f"{name}: missing gradient"
        if name in skip_bias_before_bn:
            # a bias feeding train-mode BatchNorm has an exactly-zero gradient (the batch mean is
            # subtracted); the reference leaves fp32 rounding noise there, so compare magnitudes
            scale = max(n.grad.abs().max().item() for n in named.values() if n.grad is not None)
            assert g.abs().max().item() < 1e-6 * scale, f"{name}: should be ~0"
            continue
        r = _rel(g, gr)
        worst = max(worst, r)
        print(f"[{tag}] grad {name:28s} l2-rel {r:.3e}")
        # near-zero gradients (parameters whose effect a later standardisation cancels) are compared
        # on the scale of the largest gradient of the network
        t = tol(name) if callable(tol) else tol
        if (g.detach().float().cpu() - gr.float()).norm().item() >= t * gr.float().norm().item() + floor * gmax:
            bad.append(f"{name}: {r:.3e}")
    assert not bad, f"[{tag}] gradients beyond l2-rel {tol}: {bad}"
    return worst


TOL = 2e-2          # north_star: bf16 tensor-core paths; the emulated-fp32 trunks measure ~1e-4
TOL_OUT = 1e-3      # outputs of the emulated-fp32 trunks: the fp32 tolerance of north_star


def _pn2_tol(name):
    """bf16 mode only: 2e-2 for everything downstream of the last max-pool (BatchNorm 13, final_fc); the
    layers behind the pools see the arg-max re-routing described in the module docstring."""
    return 2e-2 if (name.startswith("final_fc") or name.startswith("video_conv.13")) else 1.5e-1


def _pn1_tol(name):
    """bf16 mode only: every PolicyNetwork1UNet gradient passes the two 2x2 max-pools on the 3- and
    1-channel maps (rovr/policy_net_1.py:80-82) and the eps-free standardisation (:91-93); with bf16
    operands the bf16-storage ORACLE ITSELF moves by 0.22-0.44 L2-rel when the input is scaled by
    1 + 1e-6. The bound only catches gross errors."""
    return 6e-2 if name.startswith("fc_final") else 6e-1


_PN1_BIAS = tuple(f"{c}.bias" for c in ["conv1", "conv2", "conv3", "conv4", "upconv1", "conv5", "upconv2", "conv6",
                                        "upconv3", "conv7", "conv8", "conv9"])
_PN2_BIAS = tuple(f"video_conv.{i}.bias" for i in (0, 4, 8, 12))


def test_pn1_state_dict_layout(golden_dir):
    from policy_net_1 import PolicyNetwork1UNet
    G = _golden(golden_dir, "pn1.npz")
    net = PolicyNetwork1UNet()
    assert [n for n, _ in net.named_parameters()] == list(G["param_order"])
    assert list(net.state_dict().keys()) == list(O.pn1_state_dict(0).keys())


def test_pn1_logprob_and_critic(golden_dir, monkeypatch):
    """b = 25 (the config's batch, rovr/rovr.py T = 25 frames): logprob + critic, outputs and EVERY
    parameter gradient within the north_star tolerance of the fp32 oracle; the b = 5 golden of the
    unmodified reference pins the oracle."""
    import policy_net_1 as M
    dev = _dev()
    G = _golden(golden_dir, "pn1.npz")
    sd = O.pn1_state_dict(0, False)
    # the golden (b = 5) pins the oracle to the unmodified reference
    image5, context5, action5 = _pn1_inputs(5, 11)
    torch.manual_seed(777)
    expo5 = torch.empty((5, 25)).exponential_()
    lp5 = O.pn1_logprob(sd, image5, context5, action5, expo5)
    assert np.allclose(lp5.numpy(), G["actor/logprob"], rtol=2e-4, atol=1e-5)
    net = M.PolicyNetwork1UNet(is_critic=False)
    net.load_state_dict(sd, strict=True)
    net = net.to(dev).train()
    assert net.trunk_precision == "fp32x"
    _fixed_noise(monkeypatch, M, expo5)
    lp = net.logprob(image5.to(dev), context5.to(dev), action5.to(dev))
    _close("pn1.logprob(b=5) vs reference golden", lp, torch.from_numpy(G["actor/logprob"]), TOL_OUT)
    for n, bfr in net.named_buffers():          # running statistics of every BatchNorm follow nn.BatchNorm2d
        ref = torch.from_numpy(G[f"actor/buf/{n}"])
        if bfr.dtype == torch.long:
            assert int(bfr) == int(ref), n
        else:
            assert _rel(bfr, ref) < 1e-4, f"buffer {n}: {_rel(bfr, ref):.3e}"
    # b = 25
    b = 25
    image, context, action = _pn1_inputs(b, 13)
    torch.manual_seed(781)
    expo = torch.empty((b, 25)).exponential_()
    _fixed_noise(monkeypatch, M, expo)
    net.load_state_dict(sd, strict=True)
    net.zero_grad()
    lp = net.logprob(image.to(dev), context.to(dev), action.to(dev))
    lp.sum().backward()
    leaf = _leaf(sd)
    lp_ref = O.pn1_logprob(leaf, image, context, action, expo)
    lp_ref.sum().backward()
    _close("pn1.logprob(b=25)", lp, lp_ref, TOL_OUT)
    worst = _check_param_grads("pn1.actor vs fp32 oracle", net, {k: v.grad for k, v in leaf.items() if v.requires_grad},
                               TOL, _PN1_BIAS)
    print(f"pn1.actor worst gradient l2-rel vs fp32 oracle: {worst:.3e}")
    # critic
    sdc = O.pn1_state_dict(0, True)
    crit = M.PolicyNetwork1UNet(is_critic=True)
    crit.load_state_dict(sdc, strict=True)
    crit = crit.to(dev).train()
    v5 = crit(image5.to(dev), context5.to(dev))
    _close("pn1.critic(b=5) vs reference golden", v5, torch.from_numpy(G["critic/value"]), TOL_OUT)
    crit.load_state_dict(sdc, strict=True)
    crit.zero_grad()
    v = crit(image.to(dev), context.to(dev))
    (v ** 2).sum().backward()
    leafc = _leaf(sdc)
    v_ref = O.pn1_forward(leafc, image, context, True)
    (v_ref ** 2).sum().backward()
    _close("pn1.critic(b=25) vs fp32 oracle", v, v_ref, TOL_OUT)
    assert (v.cpu() - v_ref.detach()).abs().max().item() < TOL * max(1.0, v_ref.abs().max().item())
    worst = _check_param_grads("pn1.critic vs fp32 oracle", crit, {k: t.grad for k, t in leafc.items() if t.requires_grad},
                               TOL, _PN1_BIAS)
    print(f"pn1.critic worst gradient l2-rel vs fp32 oracle: {worst:.3e}")


def test_pn1_actor_forward_index(golden_dir, monkeypatch):
    import policy_net_1 as M
    dev = _dev()
    G = _golden(golden_dir, "pn1.npz")
    sd = O.pn1_state_dict(0, False)
    net = M.PolicyNetwork1UNet(is_critic=False)
    net.load_state_dict(sd, strict=True)
    net = net.to(dev).train()
    image1, context1, _ = _pn1_inputs(1, 12)
    torch.manual_seed(778)
    expo1 = torch.empty((1, 25)).exponential_()
    _fixed_noise(monkeypatch, M, expo1)
    idx, logp = net(image1.to(dev), context1.to(dev))
    assert idx.dtype == torch.int64 and not idx.requires_grad and not logp.requires_grad
    print("pn1 fwd idx", idx.tolist(), "golden", G["actor/fwd_idx"].tolist(), "logp", logp.tolist(), G["actor/fwd_logp"].tolist())
    assert np.array_equal(idx.cpu().numpy(), G["actor/fwd_idx"]), "selected frame index differs from the reference"
    assert abs(float(logp) - float(G["actor/fwd_logp"][0])) < 5e-2
    with pytest.raises(Exception):
        M.PolicyNetwork1UNet(is_critic=True).logprob(image1, context1, torch.zeros(1, dtype=torch.long))


def test_pn2_state_dict_layout(golden_dir):
    from policy_net_2 import PolicyNetwork2UNet
    G = _golden(golden_dir, "pn2.npz")
    net = PolicyNetwork2UNet()
    assert [n for n, _ in net.named_parameters()] == list(G["param_order"])
    assert list(net.state_dict().keys()) == list(O.pn2_state_dict(0).keys())


def test_pn2_il_logprob_critic(golden_dir, monkeypatch):
    """Imitation-learning logits, PPO logprob and critic at b = 20 (the config's clip length): outputs and
    EVERY parameter gradient within the north_star tolerance of the fp32 oracle, which the goldens pin to
    the unmodified reference."""
    import policy_net_2 as M
    dev = _dev()
    G = _golden(golden_dir, "pn2.npz")
    sd = O.pn2_state_dict(0, False)
    enc, feat, target, action = _pn2_inputs(20, 21)
    net = M.PolicyNetwork2UNet(is_critic=False)
    net.load_state_dict(sd, strict=True)
    net = net.to(dev).train()
    assert net.trunk_precision == "fp32x"
    # imitation-learning entry: extra=True -> masked, standardised logits [20, 20]
    logits = net(enc.to(dev), feat.to(dev), target.to(dev), extra=True)
    (logits ** 2).sum().backward()
    leaf = _leaf(sd)
    ref = O.pn2_forward(leaf, enc, feat, target, False, extra=True)
    (ref ** 2).sum().backward()
    assert np.allclose(ref.detach().numpy(), G["actor/il_logits"], rtol=5e-4, atol=5e-5)
    _close("pn2.il_logits vs reference golden", logits, torch.from_numpy(G["actor/il_logits"]), TOL_OUT)
    worst = _check_param_grads("pn2.il vs fp32 oracle", net, {k: v.grad for k, v in leaf.items() if v.requires_grad},
                               TOL, _PN2_BIAS)
    print(f"pn2.il worst gradient l2-rel vs fp32 oracle: {worst:.3e}")
    for n, bfr in net.named_buffers():
        refb = torch.from_numpy(G[f"actor/il/buf/{n}"])
        if bfr.dtype == torch.long:
            assert int(bfr) == int(refb), n
        else:
            assert _rel(bfr, refb) < 1e-4, f"buffer {n}: {_rel(bfr, refb):.3e}"
    # PPO logprob
    net.load_state_dict(sd, strict=True)
    net.zero_grad()
    torch.manual_seed(779)
    expo = torch.empty((20, 20)).exponential_()
    _fixed_noise(monkeypatch, M, expo)
    lp = net.logprob(enc[:, 0].to(dev), feat[:, 0].to(dev), target[:, 0].to(dev), action.to(dev), dev)
    lp.sum().backward()
    leaf = _leaf(sd)
    lp_ref = O.pn2_logprob(leaf, enc[:, 0], feat[:, 0], target[:, 0], action, expo)
    lp_ref.sum().backward()
    assert np.allclose(lp_ref.detach().numpy(), G["actor/logprob"], rtol=5e-4, atol=5e-5)
    _close("pn2.logprob vs reference golden", lp, torch.from_numpy(G["actor/logprob"]), TOL_OUT)
    worst = _check_param_grads("pn2.lp vs fp32 oracle", net, {k: v.grad for k, v in leaf.items() if v.requires_grad},
                               TOL, _PN2_BIAS)
    print(f"pn2.logprob worst gradient l2-rel vs fp32 oracle: {worst:.3e}")
    # critic
    sdc = O.pn2_state_dict(0, True)
    crit = M.PolicyNetwork2UNet(is_critic=True)
    crit.load_state_dict(sdc, strict=True)
    crit = crit.to(dev).train()
    v = crit(enc[:, 0].to(dev), feat[:, 0].to(dev), target[:, 0].to(dev))
    (v ** 2).sum().backward()
    leafc = _leaf(sdc)
    v_ref = O.pn2_forward(leafc, enc[:, 0], feat[:, 0], target[:, 0], True)
    (v_ref ** 2).sum().backward()
    assert np.allclose(v_ref.detach().numpy(), G["critic/value"], rtol=5e-4, atol=5e-5)
    # the critic divides every feature by (its std over the batch + .001) (rovr/policy_net_2.py:104-106), which
    # amplifies trunk rounding: the value must still meet the tolerance against the reference golden
    _close("pn2.critic vs reference golden", v, torch.from_numpy(G["critic/value"]), TOL)
    worst = _check_param_grads("pn2.critic vs fp32 oracle", crit, {k: t.grad for k, t in leafc.items() if t.requires_grad},
                               TOL, _PN2_BIAS)
    print(f"pn2.critic worst gradient l2-rel vs fp32 oracle: {worst:.3e}")


def test_policy_bf16_mode_bounds(monkeypatch):
    """`trunk_precision = "bf16"` (opt-in): plain bf16-operand trunks. Forward values within 2e-2 of fp32;
    gradients are compared with the oracle's bf16-STORAGE emulation (which shows the same arg-max
    re-routing without any CUDA code) at the documented bounds."""
    import policy_net_1 as M1
    import policy_net_2 as M2
    dev = _dev()
    sd = O.pn2_state_dict(0, False)
    enc, feat, target, _ = _pn2_inputs(20, 21)
    net = M2.PolicyNetwork2UNet(is_critic=False)
    net.trunk_precision = "bf16"
    net.load_state_dict(sd, strict=True)
    net = net.to(dev).train()
    logits = net(enc.to(dev), feat.to(dev), target.to(dev), extra=True)
    (logits ** 2).sum().backward()
    _close("pn2(bf16).il_logits vs fp32 oracle", logits, O.pn2_forward(sd, enc, feat, target, False, extra=True), 2e-2)
    emu = _leaf(sd)
    (O.pn2_forward(emu, enc, feat, target, False, extra=True, bf16=True) ** 2).sum().backward()
    _check_param_grads("pn2(bf16) vs bf16-storage oracle", net, {k: v.grad for k, v in emu.items() if v.requires_grad},
                       _pn2_tol, _PN2_BIAS, floor=2e-2)
    sd1 = O.pn1_state_dict(0, True)
    image, context, _ = _pn1_inputs(5, 11)
    crit = M1.PolicyNetwork1UNet(is_critic=True)
    crit.trunk_precision = "bf16"
    crit.load_state_dict(sd1, strict=True)
    crit = crit.to(dev).train()
    v = crit(image.to(dev), context.to(dev))
    (v ** 2).sum().backward()
    v_emu = O.pn1_forward(sd1, image, context, True, bf16=True)
    assert (v.cpu() - v_emu.detach()).abs().max().item() < 5e-2 * max(1.0, v_emu.abs().max().item())
    emuc = _leaf(sd1)
    (O.pn1_forward(emuc, image, context, True, bf16=True) ** 2).sum().backward()
    _check_param_grads("pn1(bf16).critic vs bf16-storage oracle", crit,
                       {k: t.grad for k, t in emuc.items() if t.requires_grad}, _pn1_tol, _PN1_BIAS, floor=2e-2)


def test_pn2_actor_forward_indices(golden_dir, monkeypatch):
    import policy_net_2 as M
    dev = _dev()
    G = _golden(golden_dir, "pn2.npz")
    sd = O.pn2_state_dict(0, False)
    enc, feat, target, _ = _pn2_inputs(20, 21)
    net = M.PolicyNetwork2UNet(is_critic=False)
    net.load_state_dict(sd, strict=True)
    net = net.to(dev).train()
    torch.manual_seed(780)
    expo1 = torch.empty((1, 20)).exponential_()
    _fixed_noise(monkeypatch, M, expo1)
    idx, logp = net(enc[:1].to(dev), feat[:1].to(dev), target[:1].to(dev))
    print("pn2 fwd idx", idx.tolist(), "golden", G["actor/fwd_idx"].tolist(), "logp", logp.tolist(), G["actor/fwd_logp"].tolist())
    assert idx.shape == (1, 2) and idx.dtype == torch.int64
    assert np.array_equal(idx.cpu().numpy(), G["actor/fwd_idx"]), "selected frame indices differ from the reference"
    assert abs(float(logp) - float(G["actor/fwd_logp"][0])) < 5e-2
    with pytest.raises(Exception):
        M.PolicyNetwork2UNet(is_critic=True).get_masked_logits(None, None)
    with pytest.raises(RuntimeError):
        net(enc[:1], feat[:1], target[:1])  # CPU tensors: no fallback


def test_action_lstm_matches_reference(golden_dir):
    from action_lstm import ActionLSTM
    dev = _dev()
    G = _golden(golden_dir, "action_lstm.npz")
    m = ActionLSTM(64, 1, 2)
    assert list(m.state_dict().keys()) == list(G["keys"])
    # same deterministic weights as tests/golden/make_golden.py::block_state_dict(m, 51)
    g = torch.Generator().manual_seed(51)
    sd = {}
    for k, v in m.state_dict().items():
        if v.dim() >= 2:
            sd[k] = (torch.rand(v.shape, generator=g) - 0.5) * 2.0 * (3.0 / v.shape[-1]) ** 0.5
        else:
            sd[k] = (torch.rand(v.shape, generator=g) - 0.5) * 0.2
    m.load_state_dict(sd, strict=True)
    m = m.to(dev)
    g = torch.Generator().manual_seed(52)
    hx, cx = torch.zeros(2, 64), torch.zeros(2, 64)
    for step in range(2):
        action = torch.randint(0, 48, (2, 3), generator=g)
        new_tensor = torch.rand((2, 3, 3, 16, 16), generator=g)
        y = m(action.to(dev), new_tensor.to(dev))
        assert y.shape == (2, 3, 80, 80)
        _close(f"lstm.step{step}.y", y, torch.from_numpy(G[f"step{step}/y"]), 1e-4)
        _close(f"lstm.step{step}.hx", m.hx, torch.from_numpy(G[f"step{step}/hx"]), 1e-4)
        y_ref, hx, cx = O.action_lstm_step(sd, action, new_tensor, hx, cx)
        _close(f"lstm.step{step}.oracle", y, y_ref, 1e-4)
    # gradients flow through both steps (hidden state carries the graph, as in the reference)
    m.zero_grad()
    (y ** 2).sum().backward()
    leaf = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    g = torch.Generator().manual_seed(52)
    hx, cx = torch.zeros(2, 64), torch.zeros(2, 64)
    for step in range(2):
        action = torch.randint(0, 48, (2, 3), generator=g)
        new_tensor = torch.rand((2, 3, 3, 16, 16), generator=g)
        y_ref, hx, cx = O.action_lstm_step(leaf, action, new_tensor, hx, cx)
    (y_ref ** 2).sum().backward()
    for n, p in m.named_parameters():
        _close(f"lstm.grad.{n}", p.grad, leaf[n].grad, 1e-3)
    m.reset_hidden_states()
    assert float(m.hx.abs().sum()) == 0.0


def test_graphed_function_matches_eager():
    """graphs.GraphedFunction: the PN2 imitation-learning step replayed from a CUDA graph gives the
    eager logits and gradients bit for bit, for the captured inputs and for new ones."""
    import policy_net_2 as M
    from graphs import GraphedFunction
    dev = _dev()
    sd = O.pn2_state_dict(0, False)
    net = M.PolicyNetwork2UNet(is_critic=False)
    net.load_state_dict(sd, strict=True)
    net = net.to(dev).train()
    enc, feat, target, _ = _pn2_inputs(20, 21)
    enc2, feat2, target2, _ = _pn2_inputs(20, 22)

    def eager(e, f, t):
        net.zero_grad(set_to_none=True)
        out = net(e.to(dev), f.to(dev), t.to(dev), extra=True)
        (out ** 2).sum().backward()
        return out.detach().clone(), {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}

    o1, g1 = eager(enc, feat, target)
    o2, g2 = eager(enc2, feat2, target2)
    net.zero_grad(set_to_none=True)
    box = {}

    def fn(e, f, t):
        box["out"] = net(e, f, t, extra=True)
        (box["out"] ** 2).sum().backward()
        return box["out"]

    step = GraphedFunction(fn, (enc.to(dev), feat.to(dev), target.to(dev)), modules=[net])
    assert step.launches > 40
    for (e, f, t), (o_ref, g_ref) in (((enc, feat, target), (o1, g1)), ((enc2, feat2, target2), (o2, g2))):
        out = step(e.to(dev), f.to(dev), t.to(dev))
        torch.cuda.synchronize()
        assert torch.equal(out.detach(), o_ref)
        for n, p in net.named_parameters():
            if n in g_ref:
                assert torch.equal(p.grad, g_ref[n]), n


def test_pn2_eval_mode_batchnorm():
    """module.eval(): BatchNorm uses (and does not touch) the running statistics; forward and gradients
    follow the oracle evaluated the same way (nn.BatchNorm2d eval semantics)."""
    import policy_net_2 as M
    dev = _dev()
    sd = O.pn2_state_dict(0, False)
    g = torch.Generator().manual_seed(5)
    for k in list(sd.keys()):
        if k.endswith("running_mean"):
            sd[k] = 0.3 * torch.randn(sd[k].shape, generator=g)
        if k.endswith("running_var"):
            sd[k] = 0.5 + torch.rand(sd[k].shape, generator=g)
    enc, feat, target, _ = _pn2_inputs(20, 23)
    net = M.PolicyNetwork2UNet(is_critic=False)
    net.load_state_dict(sd, strict=True)
    net = net.to(dev).eval()
    logits = net(enc.to(dev), feat.to(dev), target.to(dev), extra=True)
    (logits ** 2).sum().backward()
    for n, bfr in net.named_buffers():
        assert torch.equal(bfr.cpu(), sd[n]), f"eval mode must not update {n}"
    with O.bn_eval_mode():
        leaf = _leaf(sd)
        ref = O.pn2_forward(leaf, enc, feat, target, False, extra=True)
        (ref ** 2).sum().backward()
    _close("pn2.eval.il_logits", logits, ref, TOL_OUT)
    grads = {k: v.grad for k, v in leaf.items() if v.requires_grad}
    _check_param_grads("pn2.eval vs fp32 oracle", net, grads, TOL)
    # in eval mode the convolution biases DO receive a gradient (the statistics are constants)
    assert net.video_conv[0].bias.grad.abs().max().item() > 0
