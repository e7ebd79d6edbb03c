"""GPU operator tests of the emulated-fp32 path (csrc/fp32x.cuh + the fp32-output igemm epilogues):
every kernel through the C ABI against PyTorch (fp64 where the claim is fp32-class accuracy).

The path exists because the policy trunks' gradients pass through max-pools whose arg-max needs
fp32-class convolution outputs: 6-term split-bf16 forward products must match an
fp64 convolution to ~1e-6, 3-term gradient products to ~1e-4."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def _dev():
    import _native
    _native.require_device()
    return torch.device("cuda:0")


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-300)).item()


def _nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def _nchw(t):
    return t.permute(0, 3, 1, 2).contiguous()


def _pieces(x):
    h = x.to(BF).float()
    m = (x - h).to(BF).float()
    l = (x - h - m).to(BF).float()
    return h, m, l


@pytest.mark.parametrize("nterms", [6, 3, 2])
def test_split_stack_pieces_and_pool(nterms):
    import ops
    dev = _dev()
    g = torch.Generator().manual_seed(1)
    x = (torch.randn((2, 12, 16, 20), generator=g) * 3).to(dev)           # NHWC, C = 20 -> cb = 24
    s = ops.split_stack(x, nterms, 24)
    assert s.shape == (2, 12, 16, nterms * 24) and s.dtype == BF
    h, m, l = _pieces(x)
    want = {6: [h, m, l, h, m, h], 3: [h, m, h], 2: [h, m]}[nterms]
    for t, w in enumerate(want):
        blk = s[..., t * 24:(t + 1) * 24].float()
        assert torch.equal(blk[..., :20], w), (nterms, t)
        assert torch.count_nonzero(blk[..., 20:]) == 0
    if nterms == 6:
        assert torch.equal((h + m) + l, x)                               # three bf16 pieces carry all 24 bits
    # channel-slice source + fused 2x2 / (2, stride (2,1)) pooling
    xs = x[..., 4:12]
    for pool in [(2, 2, 2, 2), (4, 4, 4, 4), (2, 2, 2, 1)]:
        sp = ops.split_stack(xs, nterms, 8, pool=pool)
        ref = _nhwc(F.max_pool2d(_nchw(xs), (pool[0], pool[1]), (pool[2], pool[3])))
        assert torch.equal(sp[..., :8].float(), ref.to(BF).float()), pool


def test_split_stack_nchw_two_sources():
    import ops
    dev = _dev()
    g = torch.Generator().manual_seed(2)
    a = torch.rand((3, 3, 16, 24), generator=g).to(dev)
    b = torch.rand((3, 3, 16, 24), generator=g).to(dev)
    s = ops.split_stack(a, 6, 8, layout="nchw", src2=b)
    cat = _nhwc(torch.cat([a, b], 1))
    h, m, l = _pieces(cat)
    assert torch.equal(s[..., 0:6].float(), h) and torch.equal(s[..., 8:14].float(), m) and torch.equal(s[..., 16:22].float(), l)
    assert torch.count_nonzero(s[..., 6:8]) == 0
    one = ops.split_stack(a[:, :1].contiguous(), 6, 8, layout="nchw")
    assert torch.equal(one[..., 0].float(), a[:, 0].to(BF).float())


@pytest.mark.parametrize("shape", [(2, 24, 40, 6, 32), (1, 16, 8, 64, 128), (3, 5, 5, 256, 64), (2, 20, 20, 32, 64)])
def test_conv3x3_split_fprop_dgrad_wgrad(shape):
    """6-term forward, 3-term data gradient, 4-block weight gradient vs an fp64 convolution."""
    import ops
    from _blocks import pad16
    dev = _dev()
    B, H, W, Cin, Cout = shape
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn((B, Cin, H, W), generator=g).to(dev)
    w = (torch.randn((Cout, Cin, 3, 3), generator=g) / (3 * Cin ** 0.5)).to(dev)
    bias = torch.randn(Cout, generator=g).to(dev)
    gy = torch.randn((B, Cout, H, W), generator=g).to(dev)
    xd, wd_, bd = x.double().requires_grad_(True), w.double().requires_grad_(True), bias.double()
    yref = F.conv2d(xd, wd_, bd, padding=1)
    gxref, gwref = torch.autograd.grad(yref, (xd, wd_), gy.double())
    cbi, cbo = ops.pad8(Cin), pad16(Cout)
    xs = ops.split_stack(_nhwc(x), 6, cbi)
    wk = ops.repack_conv3x3(ops.split_weights(w, 1, 6, cbi), False)
    y = torch.empty((B, H, W, Cout), dtype=torch.float32, device=dev)
    ops.conv3x3_f32out(xs, wk, bias, y)
    e_f = _rel(_nchw(y), yref)
    # plain bf16 operands for comparison (what the emulation buys)
    e_bf = _rel(F.conv2d(x.to(BF).float(), w.to(BF).float(), bias, padding=1), yref)
    print(f"conv3x3 {shape}: fwd 6-term {e_f:.2e} (bf16 operands {e_bf:.2e})")
    assert e_f < 2e-5 and e_f < e_bf / 100       # fp32-class: what a cuDNN fp32 convolution shows at this K
    ds = ops.split_stack(_nhwc(gy), 3, cbo)
    wdk = ops.repack_conv3x3(ops.split_weights(w, 0, 3, cbo), True)
    gx = torch.empty((B, H, W, pad16(Cin)), dtype=torch.float32, device=dev)
    ops.conv3x3_f32out(ds, wdk, None, gx)
    e_d = _rel(_nchw(gx[..., :Cin]), gxref)
    assert torch.count_nonzero(gx[..., Cin:]) == 0
    dwp = torch.empty((2 * cbo, 2 * cbi, 3, 3), dtype=torch.float32, device=dev)
    ops.conv3x3_wgrad(ds[..., :2 * cbo], xs[..., :2 * cbi], dwp)
    dw = torch.empty_like(w)
    ops.blocksum4(dwp, dw, cbo, cbi)
    e_w = _rel(dw, gwref)
    print(f"conv3x3 {shape}: dgrad 3-term {e_d:.2e}, wgrad 4-block {e_w:.2e}")
    assert e_d < 1e-4 and e_w < 1e-4


@pytest.mark.parametrize("shape", [(2, 10, 10, 256, 128), (1, 8, 12, 64, 32), (2, 40, 40, 64, 32)])
def test_convT2x2_split_fprop_dgrad_wgrad(shape):
    import ops
    from _blocks import pad16
    dev = _dev()
    B, H, W, Cin, Cout = shape
    g = torch.Generator().manual_seed(sum(shape) + 1)
    x = torch.randn((B, Cin, H, W), generator=g).to(dev)
    w = (torch.randn((Cin, Cout, 2, 2), generator=g) / Cin ** 0.5).to(dev)
    bias = torch.randn(Cout, generator=g).to(dev)
    gy = torch.randn((B, Cout, 2 * H, 2 * W), generator=g).to(dev)
    xd, wd_ = x.double().requires_grad_(True), w.double().requires_grad_(True)
    yref = F.conv_transpose2d(xd, wd_, bias.double(), stride=2)
    gxref, gwref = torch.autograd.grad(yref, (xd, wd_), gy.double())
    cbi, cbo = ops.pad8(Cin), pad16(Cout)
    xs = ops.split_stack(_nhwc(x), 6, cbi)
    wk = ops.repack_convT2x2(ops.split_weights(w, 0, 6, cbi), False)
    # output into a channel slice of a wider (concat) buffer, like the U-Net's [up, skip]
    cat = torch.zeros((B, 2 * H, 2 * W, 2 * Cout), dtype=torch.float32, device=dev)
    ops.convT2x2_fprop_f32out(xs, wk, bias, cat[..., :Cout])
    assert _rel(_nchw(cat[..., :Cout]), yref) < 2e-5 and torch.count_nonzero(cat[..., Cout:]) == 0
    ds = ops.split_stack(_nhwc(gy), 3, cbo)
    wdk = ops.repack_convT2x2(ops.split_weights(w, 1, 3, cbo), True)
    gx = torch.empty((B, H, W, Cin), dtype=torch.float32, device=dev)
    ops.convT2x2_dgrad_f32out(ds, wdk, gx)
    dwp = torch.empty((2 * cbi, 2 * cbo, 2, 2), dtype=torch.float32, device=dev)
    ops.convT2x2_wgrad(ds[..., :2 * cbo], xs[..., :2 * cbi], dwp)
    dw = torch.empty_like(w)
    ops.blocksum4(dwp, dw, cbi, cbo)
    e_d, e_w = _rel(_nchw(gx), gxref), _rel(dw, gwref)
    print(f"convT {shape}: dgrad {e_d:.2e} wgrad {e_w:.2e}")
    assert e_d < 1e-4 and e_w < 1e-4


@pytest.mark.parametrize("C,c_valid,train", [(64, 64, True), (16, 3, True), (16, 1, True), (128, 128, False)])
def test_bn_f32_fwd_bwd(C, c_valid, train):
    import ops
    dev = _dev()
    g = torch.Generator().manual_seed(C + c_valid)
    B, H, W = 3, 10, 14
    x = (torch.randn((B, c_valid, H, W), generator=g) * 2 + 5).to(dev)      # |mean| >> std: the centred pass matters
    gamma, beta = (torch.rand(c_valid, generator=g) + 0.5).to(dev), torch.randn(c_valid, generator=g).to(dev)
    rm, rv = torch.randn(c_valid, generator=g).to(dev), (torch.rand(c_valid, generator=g) + 0.5).to(dev)
    gy = torch.randn((B, c_valid, H, W), generator=g).to(dev)
    xr = x.double().requires_grad_(True)
    gr, br = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    rm_ref, rv_ref = rm.double().clone(), rv.double().clone()
    yr = F.relu(F.batch_norm(xr, rm_ref, rv_ref, gr, br, train, 0.1, 1e-5))
    gxr, ggr, gbr = torch.autograd.grad(yr, (xr, gr, br), gy.double())
    xp = torch.zeros((B, H, W, C), dtype=torch.float32, device=dev)
    xp[..., :c_valid] = _nhwc(x)
    yp = torch.empty_like(xp)
    nbt = torch.zeros((), dtype=torch.long, device=dev)
    if train:
        mean, rstd = ops.bn_f32_train_fwd(xp, yp, gamma, beta, c_valid, 1e-5, 0.1, rm, rv, nbt)
        assert int(nbt) == 1 and _rel(rm, rm_ref) < 1e-6 and _rel(rv, rv_ref) < 1e-5
    else:
        mean, rstd = rm, ops.bn_f32_eval_fwd(xp, yp, gamma, beta, c_valid, 1e-5, rm, rv)
    assert _rel(_nchw(yp[..., :c_valid]), yr) < 2e-6 and torch.count_nonzero(yp[..., c_valid:]) == 0
    gyp = torch.zeros_like(xp)
    gyp[..., :c_valid] = _nhwc(gy)
    dx = torch.empty_like(xp)
    dg, db = torch.empty(c_valid, device=dev), torch.empty(c_valid, device=dev)
    ops.bn_f32_bwd(gyp, yp, xp, dx, gamma, mean, rstd, c_valid, dg, db, eval_mode=not train)
    assert _rel(_nchw(dx[..., :c_valid]), gxr) < 2e-5, _rel(_nchw(dx[..., :c_valid]), gxr)
    assert _rel(dg, ggr) < 2e-5 and _rel(db, gbr) < 2e-5
    cs = torch.empty(C, device=dev)
    ops.colsum_f32(dx, cs)
    assert _rel(cs[:c_valid], gxr.sum((0, 2, 3))) < 1e-3 or float(gxr.sum((0, 2, 3)).abs().max()) < 1e-9 * float(gxr.abs().sum())


@pytest.mark.parametrize("k,s,hw", [(2, 2, (8, 12)), (8, 8, (16, 24)), (4, 4, (20, 20)), (2, (2, 1), (5, 5)), (2, (2, 2), (2, 4))])
def test_maxpool_f32_fwd_bwd(k, s, hw):
    import ops
    dev = _dev()
    g = torch.Generator().manual_seed(hw[0] * 7 + hw[1])
    x = torch.randn((2, 16, *hw), generator=g).to(dev)
    xr = x.clone().requires_grad_(True)
    yr = F.max_pool2d(xr, k, s)
    gp = torch.randn(yr.shape, generator=torch.Generator().manual_seed(3)).to(dev)
    gxr, = torch.autograd.grad(yr, xr, gp)
    xn = _nhwc(x)
    y = ops.maxpool_f32_fwd(xn, k, s)
    assert torch.equal(_nchw(y), yr.detach())
    gx = ops.maxpool_f32_bwd(xn, _nhwc(gp), torch.empty_like(xn), k, s)
    assert torch.equal(_nchw(gx), gxr)
    if isinstance(s, int) and s == k:
        skip = torch.randn(xn.shape, generator=torch.Generator().manual_seed(4)).to(dev)
        gx2 = ops.maxpool_f32_bwd(xn, _nhwc(gp), torch.empty_like(xn), k, s, gskip=skip)
        assert torch.equal(gx2, _nhwc(gxr) + skip)


def test_flatten_unflatten_f32():
    import ops
    dev = _dev()
    x = torch.randn((3, 2, 5, 16), generator=torch.Generator().manual_seed(9)).to(dev)
    rows = torch.zeros((3, 7 + 10 * 12), device=dev)
    ops.flatten_f32(x, 12, rows, 7)
    assert torch.equal(rows[:, 7:], _nchw(x)[:, :12].flatten(1)) and torch.count_nonzero(rows[:, :7]) == 0
    back = torch.full_like(x, 5.0)
    ops.unflatten_f32(rows, 12, back, 7)
    assert torch.equal(back[..., :12], x[..., :12]) and torch.count_nonzero(back[..., 12:]) == 0
