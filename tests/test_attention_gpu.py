"""GPU parity of the common_layers drop-ins (attention / feed-forward / encoder / decoder blocks,
positional encodings) against the golden fixtures produced by the unmodified reference classes
(tests/golden/make_golden.py::gen_common) and against the oracle at a larger shape.

Tolerance: bf16 tensor-core GEMMs with fp32 accumulation -> 2e-2 L2-relative on outputs, input
gradients and parameter gradients (north_star); positional encodings are fp32 (1e-5).
"""
import os

import numpy as np
import pytest
import torch

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


def _block_state_dict(module, seed):
    """tests/golden/make_golden.py::block_state_dict"""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, v in module.state_dict().items():
        if "layer_norm" in k and k.endswith("weight"):
            sd[k] = 1.0 + (torch.rand(v.shape, generator=g) - 0.5) * 0.2
        elif v.dim() >= 2:
            sd[k] = (torch.rand(v.shape, generator=g) - 0.5) * 2.0 * (3.0 / v.shape[-1]) ** 0.5
        else:
            sd[k] = (torch.rand(v.shape, generator=g) - 0.5) * 0.2
    return sd


def test_batched_gemm_and_helpers():
    import ops
    dev = _dev()
    g = torch.Generator().manual_seed(1)
    b, h, S, T, d = 2, 4, 16, 8, 32
    E = h * d
    qkv = torch.randn((b * S, 3 * E), generator=g).to(dev).to(BF)
    Q = qkv.as_strided((b, h, S, d), (S * 3 * E, d, 3 * E, 1), 0)
    K = qkv.as_strided((b, h, S, d), (S * 3 * E, d, 3 * E, 1), E)
    out = torch.empty((b, h, S, 16), dtype=torch.float32, device=dev)
    ops.gemm_batched(Q, K, out)
    ref = torch.einsum("bhsd,bhtd->bhst", Q.float(), K.float())
    assert _rel(out, ref) < 1e-5
    ob = torch.full((b * S, E + 16), 9.0, dtype=BF, device=dev)      # bf16 output into a strided head view
    Ov = ob.as_strided((b, h, S, d), (S * (E + 16), d, E + 16, 1), 0)
    P = torch.randn((b, h, S, 16), generator=g).to(dev).to(BF)
    Vt = torch.randn((b, h, d, 16), generator=g).to(dev).to(BF)
    ops.gemm_batched(P, Vt, Ov)
    ref = torch.einsum("bhst,bhdt->bhsd", P.float(), Vt.float())
    assert _rel(Ov, ref) < 1e-2 and (ob[:, E:] == 9.0).all()
    # bigger: more than one M tile and N tile
    b, h, S, d = 2, 3, 300, 64
    A = torch.randn((b, h, S, d), generator=g).to(dev).to(BF)
    Bm = torch.randn((b, h, 272, d), generator=g).to(dev).to(BF)
    out = torch.empty((b, h, S, 272), dtype=torch.float32, device=dev)
    ops.gemm_batched(A, Bm, out)
    assert _rel(out, torch.einsum("bhsd,bhtd->bhst", A.float(), Bm.float())) < 1e-5
    # transpose with padding
    x = torch.randn((2, 3, 10, 40), generator=g).to(dev).to(BF)
    t = ops.transpose_heads(x, 16)
    assert torch.equal(t[..., :10], x.transpose(2, 3)) and (t[..., 10:] == 0).all()
    # softmax fwd / bwd
    s = torch.randn((2, 3, 10, 16), generator=g).to(dev)
    p = ops.softmax_fwd(s, 13, 0.3)
    sr = s[..., :13].clone().requires_grad_(True)
    pr = torch.softmax(sr * 0.3, dim=-1)
    assert _rel(p[..., :13], pr) < 5e-3 and (p[..., 13:] == 0).all()
    dp = torch.randn((2, 3, 10, 16), generator=g).to(dev)
    ds = ops.softmax_bwd(dp, p, 13, 0.3)
    (torch.softmax(sr * 0.3, dim=-1) * dp[..., :13]).sum().backward()
    assert _rel(ds[..., :13], sr.grad) < 1e-2 and (ds[..., 13:] == 0).all()
    # GELU
    hh = torch.randn(1000, generator=g).to(dev).to(BF)
    hr = hh.float().requires_grad_(True)
    ar = torch.nn.functional.gelu(hr)
    assert _rel(ops.gelu_fwd(hh), ar) < 5e-3
    da = torch.randn(1000, generator=g).to(dev).to(BF)
    ar.backward(da.float())
    assert _rel(ops.gelu_bwd(da, hh), hr.grad) < 5e-3
    # LayerNorm
    x = torch.randn((37, 128), generator=g).to(dev) * 2 + 1
    gam, bet = torch.rand(128, generator=g).to(dev) + 0.5, torch.randn(128, generator=g).to(dev)
    yf, yb, mean, rstd = ops.layernorm_fwd(x, gam, bet, 1e-5)
    xr, gr, br = x.clone().requires_grad_(True), gam.clone().requires_grad_(True), bet.clone().requires_grad_(True)
    yr = torch.nn.functional.layer_norm(xr, (128,), gr, br, 1e-5)
    assert _rel(yf, yr) < 1e-5 and _rel(yb, yr) < 5e-3
    gy = torch.randn((37, 128), generator=g).to(dev)
    yr.backward(gy)
    dx, dg, db = torch.empty_like(x), torch.empty_like(gam), torch.empty_like(bet)
    ops.layernorm_bwd(gy, x, gam, mean, rstd, dx, dgamma=dg, dbeta=db)
    assert _rel(dx, xr.grad) < 1e-4 and _rel(dg, gr.grad) < 1e-4 and _rel(db, br.grad) < 1e-4


def _run_block(name, dev):
    import common_layers as CL
    E, heads = 128, 4
    ctor = {"self_attn": lambda: CL.SelfAttentionBlock(E, heads, 0.0), "cross_attn": lambda: CL.CrossAttentionBlock(E, heads, 0.0),
            "ffn": lambda: CL.FeedForwardBlock(E, 0.0), "encoder": lambda: CL.EncoderBlock(E, heads, 0.0),
            "decoder": lambda: CL.DecoderBlock(E, heads, 0.0)}[name]
    return ctor()


@pytest.mark.parametrize("name,nin", [("self_attn", 1), ("cross_attn", 2), ("ffn", 1), ("encoder", 1), ("decoder", 2)])
def test_blocks_vs_reference_golden(golden_dir, name, nin):
    dev = _dev()
    G = np.load(os.path.join(golden_dir, "common_layers.npz"), allow_pickle=False)
    E, heads, S, T, B = 128, 4, 16, 8, 2
    g = torch.Generator().manual_seed(31)
    x = torch.randn((B, S, E), generator=g)
    enc = torch.randn((B, T, E), generator=g)
    m = _run_block(name, dev)
    sd = _block_state_dict(m, 41)
    assert list(sd.keys()) == list(G[f"{name}/keys"]), "state_dict keys differ from the reference"
    m.load_state_dict(sd, strict=True)
    m = m.to(dev).eval()
    ins = [t.clone().to(dev).requires_grad_(True) for t in ((x, enc)[:nin])]
    y = m(*ins)
    (y ** 2).sum().backward()
    r = _rel(y, torch.from_numpy(G[f"{name}/y"]))
    print(f"{name}: y l2-rel {r:.3e}")
    assert r < 2e-2
    for i, t in enumerate(ins):
        ri = _rel(t.grad, torch.from_numpy(G[f"{name}/gin{i}"]))
        print(f"{name}: grad input{i} l2-rel {ri:.3e}")
        assert ri < 2e-2
    for n, p in m.named_parameters():
        key = f"{name}/{n}"
        gn = float(G[f"gnorm/{key}"])
        if f"gfull/{key}" in G:
            rp = _rel(p.grad, torch.from_numpy(G[f"gfull/{key}"]))
        else:
            idx = torch.from_numpy(G[f"gidx/{key}"])
            rp = _rel(p.grad.cpu().reshape(-1)[idx], torch.from_numpy(G[f"gval/{key}"]))
        print(f"{name}: grad {n:45s} l2-rel {rp:.3e} (norm {gn:.3e})")
        assert rp < 2e-2, f"{name}: grad {n} l2-rel {rp:.3e}"


def test_positional_encodings_vs_reference_golden(golden_dir):
    import common_layers as CL
    dev = _dev()
    G = np.load(os.path.join(golden_dir, "common_layers.npz"), allow_pickle=False)
    g = torch.Generator().manual_seed(31)
    torch.randn((2, 16, 128), generator=g)
    torch.randn((2, 8, 128), generator=g)
    ipe = CL.ImagePositionalEncoding(4, 4, 8)
    ipe.load_state_dict(_block_state_dict(ipe, 42), strict=True)
    xi = torch.randn((2, 16, 128), generator=g)
    xi_d = xi.to(dev).requires_grad_(True)
    yi = ipe.to(dev)(xi_d)
    assert _rel(yi, torch.from_numpy(G["ipe/y"])) < 1e-5
    cpe = CL.ContextPositionalEncoding(2, 4, 8, 2)
    cpe.load_state_dict(_block_state_dict(cpe, 43), strict=True)
    xc = torch.randn((2, 8, 128), generator=g)
    yc = cpe.to(dev)(xc.to(dev))
    assert _rel(yc, torch.from_numpy(G["cpe/y"])) < 1e-5
    # gradients of the Linear(1, D) tables against autograd on the oracle expression
    w = torch.randn((2, 16, 128), generator=g).to(dev)
    (yi * w).sum().backward()
    sd = {("p." + k): v.detach().cpu().clone().requires_grad_(True) for k, v in ipe.state_dict().items()}
    (O.image_positional_encoding(sd, "p.", xi, 4) * w.cpu()).sum().backward()
    assert _rel(ipe.positional_encoder.weight.grad, sd["p.positional_encoder.weight"].grad) < 1e-4
    assert _rel(ipe.positional_encoder.bias.grad, sd["p.positional_encoder.bias"].grad) < 1e-4
    assert _rel(xi_d.grad, w) < 1e-6


def test_encoder_block_realistic_shape_vs_oracle():
    """256 image tokens x E = 768 (the survey's 3072 scaled to keep the CPU oracle fast), 8 heads."""
    import common_layers as CL
    dev = _dev()
    E, heads, S, B = 768, 8, 256, 3
    m = CL.EncoderBlock(E, heads, 0.0)
    sd = _block_state_dict(m, 77)
    m.load_state_dict(sd, strict=True)
    m = m.to(dev).eval()
    g = torch.Generator().manual_seed(78)
    x = torch.randn((B, S, E), generator=g)
    xd = x.to(dev).requires_grad_(True)
    y = m(xd)
    (y ** 2).sum().backward()
    leaf = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xr = x.clone().requires_grad_(True)
    yr = O.encoder_block(leaf, "", xr, heads)
    (yr ** 2).sum().backward()
    print("encoder768: y", _rel(y, yr), "gx", _rel(xd.grad, xr.grad))
    assert _rel(y, yr) < 2e-2 and _rel(xd.grad, xr.grad) < 2e-2
    for n, p in m.named_parameters():
        r = _rel(p.grad, leaf[n].grad)
        print(f"encoder768: grad {n:45s} l2-rel {r:.3e}")
        assert r < 2e-2, n


def test_encoder_decoder_at_surveyed_width_3072():
    """The token shape the reference's comments give (rovr/common_layers.py:8-9,28-29): E = 3072, 256 image tokens,
    128 context tokens; EncoderBlock and DecoderBlock forward + every gradient vs the oracle, which is evaluated by
    PyTorch fp32 on the GPU (TF32 off, tests/conftest.py) because 28 M parameters x 256 tokens are slow on the host."""
    import common_layers as CL
    dev = _dev()
    E, heads, S, T, B = 3072, 8, 256, 128, 2
    g = torch.Generator().manual_seed(90)
    x = torch.randn((B, S, E), generator=g)
    enc = torch.randn((B, T, E), generator=g)
    for kind in ("encoder", "decoder"):
        m = CL.EncoderBlock(E, heads, 0.0) if kind == "encoder" else CL.DecoderBlock(E, heads, 0.0)
        sd = _block_state_dict(m, 91)
        m.load_state_dict(sd, strict=True)
        m = m.to(dev).eval()
        xd, ed = x.to(dev).requires_grad_(True), enc.to(dev).requires_grad_(True)
        y = m(xd) if kind == "encoder" else m(xd, ed)
        (y ** 2).mean().backward()
        leaf = {k: v.to(dev).clone().requires_grad_(True) for k, v in sd.items()}
        xr, er = x.to(dev).requires_grad_(True), enc.to(dev).requires_grad_(True)
        yr = O.encoder_block(leaf, "", xr, heads) if kind == "encoder" else O.decoder_block(leaf, "", xr, er, heads)
        (yr ** 2).mean().backward()
        worst = max(_rel(p.grad, leaf[n].grad) for n, p in m.named_parameters())
        print(f"{kind}3072: y {_rel(y, yr):.3e} gx {_rel(xd.grad, xr.grad):.3e} worst param grad {worst:.3e}")
        assert _rel(y, yr) < 2e-2 and _rel(xd.grad, xr.grad) < 2e-2
        if kind == "decoder":
            assert _rel(ed.grad, er.grad) < 2e-2
        for n, p in m.named_parameters():
            assert _rel(p.grad, leaf[n].grad) < 2e-2, (kind, n, _rel(p.grad, leaf[n].grad))
        del m, leaf
        torch.cuda.empty_cache()
