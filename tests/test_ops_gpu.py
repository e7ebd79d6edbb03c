"""GPU parity tests, one per C-ABI operator: the CUDA kernels (called through ctypes) against a
plain PyTorch fp32 evaluation of the same op on the same bf16-rounded inputs.

Tolerances: operands are bf16 (exactly representable in fp32), accumulation is fp32 on both
sides, so the only differences are summation order and the final bf16 rounding of the output:
   bf16 outputs : |err| <= 1e-2 * max|ref|   (bf16 ulp is 2^-8 relative; north_star allows 2e-2)
   fp32 outputs : |err| <= 2e-3 * max|ref|   (weight / bias gradients, fp32 partial sums)
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

BF = torch.bfloat16


def _dev():
    import _native
    _native.require_device()
    return torch.device("cuda:0")


@pytest.fixture(params=["pair-auto", "pair-always"], autouse=True)
def pair_mode(request):
    """Every operator test runs twice: with the default CTA-pair policy and with cta_group::2 launches
    forced wherever the shape allows, so the 256-row MMA path sees the small / ragged shapes too."""
    import _native
    prev = _native.lib.rovr_set_pair_mode(2 if request.param == "pair-always" else 1)
    yield
    _native.lib.rovr_set_pair_mode(prev)


def _rand_act(shape, gen, dev, scale=1.0, relu=False):
    t = torch.randn(shape, generator=gen, device="cpu") * scale
    if relu:
        t = t.clamp_min(0)
    return t.to(dev).to(BF)


def _nchw(x_nhwc):
    return x_nhwc.float().permute(0, 3, 1, 2).contiguous()


def _nhwc(x_nchw):
    return x_nchw.permute(0, 2, 3, 1).contiguous()


def _report(name, got, ref, tol):
    got = got.float()
    ref = ref.float()
    err = (got - ref).abs()
    scale = ref.abs().max().item() + 1e-12
    bad = err > tol * scale
    msg = (f"{name}: max_err={err.max().item():.4e} scale={scale:.4e} rel={err.max().item() / scale:.3e} "
           f"bad={int(bad.sum())}/{bad.numel()} got_absmax={got.abs().max().item():.4e} "
           f"nan={int(torch.isnan(got).sum())}")
    if bad.any():
        idx = bad.nonzero()[:5].tolist()
        msg += f" first_bad={idx}"
        for i in idx[:3]:
            msg += f" got{tuple(i)}={got[tuple(i)].item():.4f} ref={ref[tuple(i)].item():.4f}"
    print(msg)
    assert not bad.any() and not torch.isnan(got).any(), msg


CONV_SHAPES = [
    # B, H, W, Cin, Cout
    (2, 16, 16, 64, 64),
    (1, 8, 128, 64, 32),
    (2, 32, 32, 16, 64),     # bk = 16  (conv1 of LocalNet: 9 -> 16 padded)
    (2, 16, 16, 32, 64),     # bk = 32  (policy net widths)
    (1, 16, 16, 128, 128),   # 2 k-chunks per tap
    (1, 8, 8, 256, 512),     # 2 N tiles
    (3, 20, 20, 64, 64),     # ragged tiles
    (5, 10, 10, 64, 128),    # patches spanning images
    (7, 5, 5, 128, 256),     # tiny maps of policy_net_2
    (2, 64, 64, 64, 16),
]


@pytest.mark.parametrize("B,H,W,Cin,Cout", CONV_SHAPES)
def test_conv3x3_fprop(B, H, W, Cin, Cout):
    import ops
    dev = _dev()
    g = torch.Generator().manual_seed(B * 1000 + H + Cin + Cout)
    x = _rand_act((B, H, W, Cin), g, dev)
    w = (torch.randn((Cout, Cin, 3, 3), generator=g) * (1.0 / (3 * Cin ** 0.5))).to(dev)
    b = torch.randn((Cout,), generator=g).to(dev) * 0.1
    wk = ops.repack_conv3x3(w)
    # write into a channel slice of a wider buffer to exercise ld / offset handling
    ybuf = torch.full((B, H, W, Cout + 16), 7.0, dtype=BF, device=dev)
    y = ybuf[..., 8:8 + Cout]
    ops.conv3x3_fprop(x, wk, b, y, relu=True)
    torch.cuda.synchronize()
    ref = F.relu(F.conv2d(_nchw(x), w.to(BF).float(), b, padding=1))
    _report(f"conv3x3_fprop{(B, H, W, Cin, Cout)}", _nchw(y), ref, 1e-2)
    assert (ybuf[..., :8] == 7.0).all() and (ybuf[..., 8 + Cout:] == 7.0).all(), "wrote outside its slice"


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 16, 16, 64, 64), (1, 32, 24, 16, 64), (3, 20, 12, 64, 128),
                                            (1, 16, 8, 128, 256), (2, 64, 64, 64, 16)])
def test_conv3x3_fprop_fused_pool(B, H, W, Cin, Cout):
    """conv + ReLU + MaxPool2d(2, 2) from one kernel (encoder of rovr/local_net.py:52-55)."""
    import ops
    dev = _dev()
    g = torch.Generator().manual_seed(B * 77 + H + Cin + Cout)
    x = _rand_act((B, H, W, Cin), g, dev)
    w = (torch.randn((Cout, Cin, 3, 3), generator=g) * (1.0 / (3 * Cin ** 0.5))).to(dev)
    b = torch.randn((Cout,), generator=g).to(dev) * 0.1
    wk = ops.repack_conv3x3(w)
    ybuf = torch.full((B, H, W, Cout + 16), 7.0, dtype=BF, device=dev)
    y = ybuf[..., 8:8 + Cout]
    pooled = torch.full((B, H // 2, W // 2, Cout), -3.0, dtype=BF, device=dev)
    ops.conv3x3_fprop(x, wk, b, y, relu=True, pooled=pooled)
    y2 = torch.empty((B, H, W, Cout), dtype=BF, device=dev)
    ops.conv3x3_fprop(x, wk, b, y2, relu=True)
    torch.cuda.synchronize()
    assert torch.equal(y, y2), "fused pooling changed the convolution output"
    assert torch.equal(_nchw(pooled), F.max_pool2d(_nchw(y2), 2)), "pooled tile differs from max_pool2d of the stored output"
    assert (ybuf[..., :8] == 7.0).all() and (ybuf[..., 8 + Cout:] == 7.0).all()


@pytest.mark.parametrize("B,H,W,Cin", [(2, 16, 16, 128), (1, 32, 24, 64), (3, 20, 12, 128), (5, 8, 8, 128), (1, 64, 64, 128)])
def test_conv3x3_fprop_fused_tail(B, H, W, Cin):
    """conv7 + ReLU + conv8 (1x1, 64 -> 3) + sigmoid + L2 loss from one kernel equals the conv kernel
    followed by the stand-alone tail kernel (same bf16-rounded y)."""
    import ops
    dev = _dev()
    g = torch.Generator().manual_seed(B * 13 + H + Cin)
    x = _rand_act((B, H, W, Cin), g, dev)
    w = (torch.randn((64, Cin, 3, 3), generator=g) * (1.0 / (3 * Cin ** 0.5))).to(dev)
    b = torch.randn((64,), generator=g).to(dev) * 0.1
    w8 = (torch.randn((3, 64, 1, 1), generator=g) * 0.2).to(dev)
    b8 = torch.randn((3,), generator=g).to(dev) * 0.1
    tgt = torch.rand((B, 3, H, W), generator=g).to(dev)
    wk = ops.repack_conv3x3(w)
    y1 = torch.empty((B, H, W, 64), dtype=BF, device=dev)
    out1, loss1 = ops.conv3x3_fprop_tail(x, wk, b, y1, w8, b8, tgt)
    y2 = torch.empty((B, H, W, 64), dtype=BF, device=dev)
    ops.conv3x3_fprop(x, wk, b, y2, relu=True)
    out2, loss2 = ops.tail_fwd(y2, w8, b8, tgt)
    torch.cuda.synchronize()
    assert torch.equal(y1, y2)
    ref = torch.sigmoid(F.conv2d(_nchw(y2), w8, b8))
    _report("fused_tail.out", out1, ref, 1e-5)
    _report("fused_tail.out vs tail kernel", out1, out2, 1e-5)
    assert abs(loss1.item() - F.mse_loss(ref, tgt).item()) < 1e-5 * max(1.0, loss2.item())
    out3, loss3 = ops.conv3x3_fprop_tail(x, wk, b, y1, w8, b8, None)
    torch.cuda.synchronize()
    assert loss3 is None and torch.equal(out3, out1)


@pytest.mark.parametrize("B,H,W,Cin,Cout", CONV_SHAPES)
def test_conv3x3_dgrad(B, H, W, Cin, Cout):
    import ops
    dev = _dev()
    g = torch.Generator().manual_seed(B * 1000 + H + Cin + Cout + 1)
    dy = _rand_act((B, H, W, Cout), g, dev)
    w = (torch.randn((Cout, Cin, 3, 3), generator=g) * (1.0 / (3 * Cout ** 0.5))).to(dev)
    xmask = _rand_act((B, H, W, Cin), g, dev, relu=True)
    wd = ops.repack_conv3x3(w, for_dgrad=True)
    dx = torch.empty((B, H, W, Cin), dtype=BF, device=dev)
    colsum = torch.full((Cin,), 123.0, device=dev)
    ops.conv3x3_dgrad(dy, wd, dx, mask=xmask, colsum=colsum)
    torch.cuda.synchronize()
    ref = F.conv_transpose2d(_nchw(dy), w.to(BF).float(), padding=1) * (_nchw(xmask) > 0)
    _report(f"conv3x3_dgrad{(B, H, W, Cin, Cout)}", _nchw(dx), ref, 1e-2)
    # fused bias gradient = column sums of the stored (bf16) gradient
    _report("conv3x3_dgrad.colsum", colsum, dx.float().sum(dim=(0, 1, 2)), 1e-3)
    half = torch.full((Cin // 2,), 5.0, device=dev)  # only the first half of the channels (upconv bias)
    ops.conv3x3_dgrad(dy, wd, dx, mask=xmask, colsum=half)
    torch.cuda.synchronize()
    _report("conv3x3_dgrad.colsum[:half]", half, dx.float().sum(dim=(0, 1, 2))[: Cin // 2], 1e-3)
    if Cin % 128 == 0:
        # mask only the first half of the channels (U-Net concat gradient: the skip half is masked later)
        dxh = torch.empty_like(dx)
        ops.conv3x3_dgrad(dy, wd, dxh, mask=xmask, mask_cols=Cin // 2, colsum=half)
        torch.cuda.synchronize()
        m = _nchw(xmask) > 0
        m[:, Cin // 2:] = True
        _report("conv3x3_dgrad.mask_cols", _nchw(dxh), F.conv_transpose2d(_nchw(dy), w.to(BF).float(), padding=1) * m, 1e-2)
        _report("conv3x3_dgrad.mask_cols.colsum", half, dxh.float().sum(dim=(0, 1, 2))[: Cin // 2], 1e-3)


@pytest.mark.parametrize("B,H,W,Cin,Cout", CONV_SHAPES)
def test_conv3x3_wgrad(B, H, W, Cin, Cout):
    import ops
    dev = _dev()
    g = torch.Generator().manual_seed(B * 1000 + H + Cin + Cout + 2)
    dy = _rand_act((B, H, W, Cout), g, dev)
    x = _rand_act((B, H, W, Cin), g, dev)
    keep = Cin if Cin != 16 else 9
    dw = torch.empty((Cout, keep, 3, 3), dtype=torch.float32, device=dev)
    ops.conv3x3_wgrad(dy, x, dw)
    torch.cuda.synchronize()
    ref = torch.nn.grad.conv2d_weight(_nchw(x), (Cout, Cin, 3, 3), _nchw(dy), padding=1)[:, :keep]
    _report(f"conv3x3_wgrad{(B, H, W, Cin, Cout)}", dw, ref, 2e-3)


CONVT_SHAPES = [
    # B, H, W (input), Cin, Cout
    (2, 8, 8, 64, 32),
    (1, 16, 16, 128, 64),
    (2, 4, 4, 512, 256),     # N = 1024: 4 N tiles (upconv1 of LocalNet)
    (3, 10, 10, 64, 32),     # policy_net_1 sizes
    (2, 5, 5, 256, 128),
    (1, 8, 64, 32, 16),
]


@pytest.mark.parametrize("B,H,W,Cin,Cout", CONVT_SHAPES)
def test_convT2x2_fprop(B, H, W, Cin, Cout):
    import ops
    dev = _dev()
    g = torch.Generator().manual_seed(B * 1000 + H + Cin + Cout + 3)
    x = _rand_act((B, H, W, Cin), g, dev)
    w = (torch.randn((Cin, Cout, 2, 2), generator=g) * (1.0 / Cin ** 0.5)).to(dev)
    b = torch.randn((Cout,), generator=g).to(dev) * 0.1
    wk = ops.repack_convT2x2(w)
    ybuf = torch.full((B, 2 * H, 2 * W, 2 * Cout), 7.0, dtype=BF, device=dev)
    y = ybuf[..., :Cout]
    ops.convT2x2_fprop(x, wk, b, y, relu=True)
    torch.cuda.synchronize()
    ref = F.relu(F.conv_transpose2d(_nchw(x), w.to(BF).float(), b, stride=2))
    _report(f"convT2x2_fprop{(B, H, W, Cin, Cout)}", _nchw(y), ref, 1e-2)
    assert (ybuf[..., Cout:] == 7.0).all(), "wrote outside its slice"


@pytest.mark.parametrize("B,H,W,Cin,Cout", CONVT_SHAPES)
def test_convT2x2_dgrad(B, H, W, Cin, Cout):
    import ops
    dev = _dev()
    g = torch.Generator().manual_seed(B * 1000 + H + Cin + Cout + 4)
    dybuf = _rand_act((B, 2 * H, 2 * W, 2 * Cout), g, dev)
    dy = dybuf[..., :Cout]
    w = (torch.randn((Cin, Cout, 2, 2), generator=g) * (1.0 / (2 * Cout ** 0.5))).to(dev)
    xmask = _rand_act((B, H, W, Cin), g, dev, relu=True)
    wd = ops.repack_convT2x2(w, for_dgrad=True)
    dx = torch.empty((B, H, W, Cin), dtype=BF, device=dev)
    colsum = torch.full((Cin,), 123.0, device=dev)
    ops.convT2x2_dgrad(dy, wd, dx, mask=xmask, colsum=colsum)
    torch.cuda.synchronize()
    ref = F.conv2d(_nchw(dy), w.to(BF).float(), stride=2) * (_nchw(xmask) > 0)
    _report(f"convT2x2_dgrad{(B, H, W, Cin, Cout)}", _nchw(dx), ref, 1e-2)
    _report("convT2x2_dgrad.colsum", colsum, dx.float().sum(dim=(0, 1, 2)), 1e-3)


@pytest.mark.parametrize("B,H,W,Cin,Cout", CONVT_SHAPES)
def test_convT2x2_wgrad(B, H, W, Cin, Cout):
    import ops
    dev = _dev()
    g = torch.Generator().manual_seed(B * 1000 + H + Cin + Cout + 5)
    dybuf = _rand_act((B, 2 * H, 2 * W, 2 * Cout), g, dev)
    dy = dybuf[..., :Cout]
    x = _rand_act((B, H, W, Cin), g, dev)
    dw = torch.empty((Cin, Cout, 2, 2), dtype=torch.float32, device=dev)
    ops.convT2x2_wgrad(dy, x, dw)
    torch.cuda.synchronize()
    # d/dW of conv_transpose2d(x, W, stride 2): dW[ci,co,ky,kx] = sum x[b,ci,y,x] dy[b,co,2y+ky,2x+kx]
    xs = _nchw(x)
    dys = _nchw(dy)
    ref = torch.stack([torch.stack([torch.einsum("bihw,bohw->io", xs, dys[:, :, ky::2, kx::2])
                                    for kx in range(2)], -1) for ky in range(2)], -2)
    _report(f"convT2x2_wgrad{(B, H, W, Cin, Cout)}", dw, ref, 2e-3)


def test_repack_batch_matches_single_calls():
    """One launch for every weight tensor of a network == the per-tensor repack calls, bit for bit
    (all four layouts, a padded input-channel count among them)."""
    import ops
    dev = _dev()
    gen = torch.Generator().manual_seed(3)
    convs = [torch.randn(64, 9, 3, 3, generator=gen).to(dev), torch.randn(128, 64, 3, 3, generator=gen).to(dev)]
    ups = [torch.randn(128, 64, 2, 2, generator=gen).to(dev), torch.randn(32, 16, 2, 2, generator=gen).to(dev)]
    req = [(w, "conv", fd) for w in convs for fd in (False, True)] + [(w, "up", fd) for w in ups for fd in (False, True)]
    got = ops.repack_batch(req)
    for (w, kind, fd), g in zip(req, got):
        ref = ops.repack_convT2x2(w, fd) if kind == "up" else ops.repack_conv3x3(w, fd)
        assert g.shape == ref.shape and torch.equal(g, ref), (kind, fd, tuple(w.shape))


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (300, 256, 128), (25, 768, 2048), (1000, 48, 32)])
def test_gemm(M, N, K):
    import ops
    dev = _dev()
    g = torch.Generator().manual_seed(M + N + K)
    x = (torch.randn((M, K), generator=g) / K ** 0.5).to(dev).to(BF)
    w = torch.randn((N, K), generator=g).to(dev).to(BF)
    b = torch.randn((N,), generator=g).to(dev)
    y = ops.gemm_bf16(x, w, b, relu=False, out_dtype=torch.float32)
    torch.cuda.synchronize()
    ref = x.float() @ w.float().t() + b
    _report(f"gemm{(M, N, K)}", y, ref, 2e-3)
    yb = ops.gemm_bf16(x, w, b, relu=True, out_dtype=BF)
    torch.cuda.synchronize()
    _report(f"gemm_bf16relu{(M, N, K)}", yb, F.relu(ref), 1e-2)


@pytest.mark.parametrize("B,H,W,C,k,s", [(2, 16, 16, 64, 2, 2), (3, 160, 160, 64, 8, 8), (2, 20, 20, 128, 4, 4),
                                          (2, 5, 5, 512, 2, (2, 1)), (2, 2, 4, 512, 2, 2), (2, 5, 5, 256, 1, 1)])
def test_maxpool(B, H, W, C, k, s):
    import ops
    dev = _dev()
    g = torch.Generator().manual_seed(B + H + C)
    xbuf = _rand_act((B, H, W, C + 8), g, dev, relu=True)
    x = xbuf[..., 8:]
    sh, sw = (s, s) if isinstance(s, int) else s
    Ho, Wo = (H - k) // sh + 1, (W - k) // sw + 1
    y = torch.empty((B, Ho, Wo, C), dtype=BF, device=dev)
    ops.maxpool_fwd(x, y, k, s)
    xr = _nchw(x).requires_grad_(True)
    yr = F.max_pool2d(xr, k, s)
    torch.cuda.synchronize()
    assert torch.equal(_nchw(y), yr.detach()), "maxpool forward must be exact"
    gp = _rand_act((B, Ho, Wo, C), g, dev)
    tiled = (k, k) == (sh, sw) and H % k == 0 and W % k == 0
    gskip = _rand_act((B, H, W, C), g, dev) if tiled else None
    gx = torch.empty((B, H, W, C), dtype=BF, device=dev)
    ops.maxpool_bwd(x, gp, gx, k, s, gskip=gskip, relu_mask=True)
    torch.cuda.synchronize()
    yr.backward(_nchw(gp))
    ref = xr.grad
    if gskip is not None:
        ref = ref + _nchw(gskip)
    ref = ref * (xr.detach() > 0)
    # bf16 ties inside a window can route the gradient to a different (equal-valued) element
    # than ATen only if two equal maxima exist; the inputs are random normals so ties are only at 0
    _report(f"maxpool_bwd{(B, H, W, C, k, s)}", _nchw(gx), ref, 1e-2)
    if tiled and C <= 256 and 256 % (C // 8) == 0:
        # fused bias gradient: column sums of gx from the same pass
        cs = torch.full((C,), 77.0, device=dev)
        gx2 = torch.empty_like(gx)
        ops.maxpool_bwd(x, gp, gx2, k, s, gskip=gskip, relu_mask=True, colsum=cs)
        torch.cuda.synchronize()
        assert torch.equal(gx2, gx)
        _report("maxpool_bwd.colsum", cs, ref.sum(dim=(0, 2, 3)), 2e-3)


def test_pack_and_unpack():
    import ops
    dev = _dev()
    g = torch.Generator().manual_seed(5)
    a = torch.rand((3, 3, 24, 40), generator=g).to(dev)
    c = torch.rand((3, 2, 3, 24, 40), generator=g).to(dev)
    p = ops.pack_nchw([a, c.reshape(3, 6, 24, 40)], 16)
    torch.cuda.synchronize()
    ref = torch.cat([a, c.reshape(3, 6, 24, 40)], 1).to(BF)
    assert torch.equal(_nchw(p)[:, :9], ref.float())
    assert (p[..., 9:] == 0).all()
    back = ops.unpack_nhwc(p, 9)
    torch.cuda.synchronize()
    assert torch.equal(back, ref.float())


@pytest.mark.parametrize("B,H,W", [(2, 32, 32), (1, 64, 40), (3, 8, 8)])
def test_tail(B, H, W):
    import ops
    dev = _dev()
    g = torch.Generator().manual_seed(B + H)
    y7 = _rand_act((B, H, W, 64), g, dev, relu=True)
    w8 = (torch.randn((3, 64, 1, 1), generator=g) * 0.2).to(dev)
    b8 = (torch.randn((3,), generator=g) * 0.1).to(dev)
    tgt = torch.rand((B, 3, H, W), generator=g).to(dev)
    out, loss = ops.tail_fwd(y7, w8, b8, tgt)
    torch.cuda.synchronize()
    y7r = _nchw(y7).requires_grad_(True)
    w8r = w8.clone().requires_grad_(True)
    b8r = b8.clone().requires_grad_(True)
    ref = torch.sigmoid(F.conv2d(y7r, w8r, b8r))
    lref = F.mse_loss(ref, tgt)
    _report("tail_fwd", out, ref.detach(), 1e-4)
    assert abs(loss.item() - lref.item()) < 1e-4 * max(1.0, abs(lref.item()))
    # backward driven by the fused loss
    g7 = torch.empty_like(y7)
    dw8 = torch.empty((3, 64), device=dev)
    db8 = torch.empty((3,), device=dev)
    gl = torch.ones((), device=dev)
    ops.tail_bwd(y7, w8, out, g7, dw8, db8, gout=None, target=tgt, mse_scale=2.0 / out.numel(), gloss=gl)
    torch.cuda.synchronize()
    lref.backward()
    _report("tail_bwd.g7", _nchw(g7), y7r.grad * (y7r.detach() > 0), 1e-2)
    _report("tail_bwd.dw8", dw8, w8r.grad.reshape(3, 64), 2e-3)
    _report("tail_bwd.db8", db8, b8r.grad, 2e-3)
    # backward driven by an explicit upstream gradient
    gout = torch.randn((B, 3, H, W), generator=g).to(dev)
    y7r.grad = None
    w8r.grad = None
    ref2 = torch.sigmoid(F.conv2d(y7r, w8r, b8r))
    ref2.backward(gout)
    ops.tail_bwd(y7, w8, out, g7, dw8, db8, gout=gout)
    torch.cuda.synchronize()
    _report("tail_bwd.g7(gout)", _nchw(g7), y7r.grad * (y7r.detach() > 0), 1e-2)
    _report("tail_bwd.dw8(gout)", dw8, w8r.grad.reshape(3, 64), 2e-3)


@pytest.mark.parametrize("npix_shape,C", [((2, 16, 16), 64), ((3, 7, 5), 256), ((1, 64, 64), 32), ((2, 4, 4), 512)])
def test_colsum(npix_shape, C):
    import ops
    dev = _dev()
    g = torch.Generator().manual_seed(C)
    B, H, W = npix_shape
    buf = _rand_act((B, H, W, C + 8), g, dev)
    x = buf[..., :C]
    out = torch.empty((C,), device=dev)
    ops.colsum(x, out)
    torch.cuda.synchronize()
    _report("colsum", out, x.float().sum(dim=(0, 1, 2)), 1e-4)


def test_corrupt_frames_matches_dataset_logic():
    """rovr/video_ds.py:62-87 (the deterministic zeroed box; frame_index already halved as in :63) restated in numpy
    vs the device kernel, at the dataset's 256x256 and at a non-square size."""
    import numpy as np
    import ops
    dev = _dev()

    def corrupt_frame(frame, frame_index):            # frame HWC, as in the reference
        h, w, _ = frame.shape
        mask = np.ones_like(frame)
        section_idx, position_idx = frame_index // 8, frame_index % 8
        start_y = section_idx * h // 3
        end_y = start_y + 100
        start_x = position_idx * w // 8
        end_x = start_x + 150
        start_x, end_x = max(0, start_x), min(w, end_x)
        start_y, end_y = max(0, start_y), min(h, end_y)
        mask[start_y:end_y, start_x:end_x, :] = 0
        return frame * mask, mask

    for H, W in ((256, 256), (120, 200)):
        g = torch.Generator().manual_seed(H)
        clean = torch.rand((25, 3, H, W), generator=g)
        idx = torch.arange(25)
        out, mask = ops.corrupt_frames(clean.to(dev), idx.to(dev), want_mask=True)
        for n in range(25):
            ref, m = corrupt_frame(clean[n].permute(1, 2, 0).numpy(), n)
            assert np.array_equal(out[n].permute(1, 2, 0).cpu().numpy(), ref), (H, W, n)
            assert np.array_equal(mask[n].permute(1, 2, 0).cpu().numpy(), m)


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 56, 56, 128, 128), (3, 28, 28, 256, 256), (5, 14, 14, 512, 512), (1, 15, 9, 64, 32),
                                            (2, 6, 6, 32, 64)])
def test_conv3x3_stride2(B, H, W, Cin, Cout):
    """3x3 / padding 1 / stride 2 (ResNet-50's down-sampling convolutions) through TMA element strides, vs F.conv2d
    on the same bf16-rounded operands — and against the stride-1 kernel sampled at the even positions."""
    import ops
    dev = _dev()
    g = torch.Generator().manual_seed(B * 31 + H + Cin + Cout)
    x = _rand_act((B, H, W, Cin), g, dev)
    w = (torch.randn((Cout, Cin, 3, 3), generator=g) * (1.0 / (3 * Cin ** 0.5))).to(dev)
    b = torch.randn((Cout,), generator=g).to(dev) * 0.1
    wk = ops.repack_conv3x3(w)
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    y = torch.full((B, Ho, Wo, Cout), 9.0, dtype=BF, device=dev)
    ops.conv3x3_fprop_s2(x, wk, b, y, relu=True)
    ref = F.relu(F.conv2d(_nchw(x.float()), w.to(BF).float(), b, stride=2, padding=1))
    assert ref.shape == (B, Cout, Ho, Wo)
    _report("conv3x3_s2", _nchw(y), ref, 1e-2)
    y1 = torch.empty((B, H, W, Cout), dtype=BF, device=dev)
    ops.conv3x3_fprop(x, wk, b, y1, relu=True)
    # the stride-1 kernel sampled at the even positions (its halo mode accumulates the taps in another order: not
    # bit-identical, but within one bf16 rounding)
    _report("conv3x3_s2 vs sub-sampled stride 1", _nchw(y), _nchw(y1[:, ::2, ::2].contiguous()).float(), 1e-2)
    # fp32 output of the same kernel (train-mode ResNet trunk): the same accumulators, unrounded
    yf = torch.full((B, Ho, Wo, Cout), 9.0, dtype=torch.float32, device=dev)
    ops.conv3x3_fprop_s2_f32out(x, wk, b, yf, relu=True)
    assert torch.equal(yf.to(BF), y)
    _report("conv3x3_s2 fp32 out", _nchw(yf), ref, 2e-5)


@pytest.mark.parametrize("frames,H,W,C,f32", [(6, 56, 56, 64, True), (6, 7, 7, 2048, True), (3, 14, 14, 1024, False), (1, 28, 28, 512, True),
                                              (40, 9, 5, 256, True)])
def test_bn_train_fwd_frames(frames, H, W, C, f32):
    """Train-mode BatchNorm with per-frame statistics (the reference's trunk at batch 1 per frame,
    rovr/resnet_extractor.py:42-47) vs a loop of F.batch_norm calls, one frame at a time, on the same buffers."""
    import ops
    dev = _dev()
    g = torch.Generator().manual_seed(frames * 7 + C)
    x = (torch.randn((frames, H, W, C), generator=g) * (0.5 + torch.rand((C,), generator=g)) + 2.0 * torch.rand((C,), generator=g))
    x = x.to(dev) if f32 else x.to(BF).to(dev)
    gamma = (0.5 + torch.rand((C,), generator=g)).to(dev)
    beta = (torch.rand((C,), generator=g) - 0.5).to(dev)
    rm, rv = torch.rand((C,), generator=g).to(dev), (1.0 + torch.rand((C,), generator=g)).to(dev)
    nbt = torch.tensor(3, dtype=torch.int64, device=dev)
    rm0, rv0 = rm.clone(), rv.clone()
    y = torch.empty((frames, H, W, C), dtype=BF, device=dev)
    mean, rstd = ops.bn_train_fwd_frames(x, y, gamma, beta, 1e-5, 0.1, rm, rv, nbt, relu=True)
    refs = []
    for f in range(frames):
        xf = x[f:f + 1].float().permute(0, 3, 1, 2)
        refs.append(F.relu(F.batch_norm(xf, rm0, rv0, gamma, beta, training=True, momentum=0.1, eps=1e-5)))
    ref = torch.cat(refs)
    assert int(nbt) == 3 + frames
    _report("bn frames y", _nchw(y), ref, 5e-3)
    _report("bn frames running_mean", rm, rm0, 1e-5)
    _report("bn frames running_var", rv, rv0, 1e-4)
    want_mean = x.float().mean(dim=(1, 2))
    _report("bn frames mean", mean, want_mean, 1e-5)
