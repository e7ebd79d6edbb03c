"""Tensor-level wrappers over the C ABI (include/rovr_b200.h).

Every function takes torch CUDA tensors, checks layout, and launches on torch's current stream.
Activations are NHWC bf16 views ([B, H, W, C] with unit channel stride, possibly a channel slice
of a wider buffer); parameters / gradients are fp32 in the reference's PyTorch layouts.
PyTorch is used for memory and streams only — no torch operator computes anything here.
"""
import ctypes

import torch

import _native as N

_vp = ctypes.c_void_p


# Device of the tensors of the call being assembled: every tensor argument goes through _ptr(), the
# stream argument is evaluated last, and _launch() switches the CUDA context if that device is not the
# thread's current one (the reference picks `torch.device(f'cuda:{id}')` without set_device,
# rovr/train_local_net_unet.py:73-75, so a model may live on a non-current GPU).
_DEV = [None]


def _ptr(t):
    if t is None:
        return _vp(0)
    if t.is_cuda:
        _DEV[0] = t.device.index
    return _vp(t.data_ptr())


def _stream():
    return _vp(torch.cuda.current_stream(_DEV[0]).cuda_stream)


# Optional per-call timing for bench.py's roofline: when a list is installed with
# `set_profile(list)`, every C-ABI call is bracketed by CUDA events on the launching stream and
# (name, start_event, end_event) is appended. None (the default) adds no overhead.
_PROFILE = None


def set_profile(sink):
    global _PROFILE
    _PROFILE = sink


def _launch(name, *args):
    dev = _DEV[0]
    if dev is not None and dev != torch.cuda.current_device():
        with torch.cuda.device(dev):
            return _launch_here(name, *args)
    return _launch_here(name, *args)


def _launch_here(name, *args):
    if _PROFILE is None:
        return N.call(name, *args)
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    N.call(name, *args)
    e1.record()
    _PROFILE.append((name, e0, e1))


def _act(t, name="activation"):
    """Validate an NHWC bf16 view and return (B, H, W, C, ld)."""
    if t.dtype != torch.bfloat16 or not t.is_cuda or t.dim() != 4:
        raise ValueError(f"{name}: expected a 4-D CUDA bf16 NHWC tensor, got {t.dtype} {tuple(t.shape)}")
    B, H, W, C = t.shape
    ld = t.stride(2)
    if t.stride(3) != 1 or t.stride(1) != W * ld or (B > 1 and t.stride(0) != H * W * ld):
        raise ValueError(f"{name}: not a dense-pixel NHWC view, strides={t.stride()}")
    return B, H, W, C, ld


def _f32(t, name):
    if t is None:
        return
    if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous():
        raise ValueError(f"{name}: expected a contiguous CUDA fp32 tensor")


# ---------------------------------------------------------------------------------------------
# scratch workspace: one per (device, stream) — ops on one stream run in order, so they can share
# it; two streams never do. A buffer that has been handed out is NEVER freed: a CUDA graph captured
# earlier has its address baked in, so when a later call needs more the old buffer is retired (kept
# alive) and a new one of at least twice the size takes over (total retired bytes < the live size).
# During a capture the capture stream is its own key, so a graph gets a workspace of its own.
# ---------------------------------------------------------------------------------------------
_WS = {}
_WS_RETIRED = []


def workspace(nbytes, device):
    index = device.index if device.index is not None else torch.cuda.current_device()
    key = (index, torch.cuda.current_stream(index).cuda_stream)
    buf = _WS.get(key)
    if buf is None or buf.numel() < nbytes:
        if buf is not None:
            _WS_RETIRED.append(buf)
        size = max(int(nbytes), 1 << 20, 2 * buf.numel() if buf is not None else 0)
        buf = torch.empty(size, dtype=torch.uint8, device=torch.device("cuda", index))
        _WS[key] = buf
    return buf


# ---------------------------------------------------------------------------------------------
# packing
# ---------------------------------------------------------------------------------------------
def pack_nchw(srcs, cpad, out=None):
    """cat(srcs, dim=1) -> NHWC bf16 with channels zero-padded to cpad. srcs: fp32 [B, c_i, H, W]."""
    srcs = [s.contiguous() for s in srcs]
    for s in srcs:
        _f32(s, "pack source")
    B, _, H, W = srcs[0].shape
    if out is None:
        out = torch.empty((B, H, W, cpad), dtype=torch.bfloat16, device=srcs[0].device)
    args = []
    for i in range(3):
        if i < len(srcs):
            args += [_ptr(srcs[i]), srcs[i].shape[1]]
        else:
            args += [_vp(0), 0]
    _launch("rovr_pack_nchw_to_nhwc", *args, _ptr(out), B, H, W, cpad, _stream())
    return out


def u8_to_f32(src, out=None, denom=255.0):
    """ToTensor on the device: uint8 tensor (any shape, contiguous) -> fp32 / denom (bit-identical to .div(255))."""
    if src.dtype != torch.uint8 or not src.is_cuda or not src.is_contiguous():
        raise ValueError("u8_to_f32: expected a contiguous CUDA uint8 tensor")
    if out is None:
        out = torch.empty(src.shape, dtype=torch.float32, device=src.device)
    assert out.numel() == src.numel() and out.is_contiguous() and out.dtype == torch.float32
    _launch("rovr_u8_to_f32", _ptr(src), _ptr(out), src.numel(), ctypes.c_float(denom), _stream())
    return out


def corrupt_frames(clean, frame_index, box_w=150, box_h=100, want_mask=False):
    """rovr/video_ds.py:62-87 on the device: (clean * mask[, mask]) for NCHW fp32 images whose clip-frame indices are
    `frame_index` (int64 [N], on the device)."""
    _f32(clean, "clean")
    N_, C, H, W = clean.shape
    fi = _idx(frame_index, "frame_index")
    assert fi.numel() == N_
    out = torch.empty_like(clean)
    mask = torch.empty_like(clean) if want_mask else None
    _launch("rovr_corrupt_frames", _ptr(clean), _ptr(fi), _ptr(out), _ptr(mask), N_, C, H, W, int(box_w), int(box_h), _stream())
    return (out, mask) if want_mask else out


def unpack_nhwc(x, C=None):
    B, H, W, Cx, ld = _act(x)
    C = Cx if C is None else C
    out = torch.empty((B, C, H, W), dtype=torch.float32, device=x.device)
    _launch("rovr_unpack_nhwc_to_nchw", _ptr(x), ld, _ptr(out), B, H, W, C, _stream())
    return out


# ---------------------------------------------------------------------------------------------
# weight repacking
# ---------------------------------------------------------------------------------------------
def pad16(c):
    return (c + 15) // 16 * 16


def repack_conv3x3(w, for_dgrad=False):
    _f32(w, "conv weight")
    Cout, Cin = w.shape[0], w.shape[1]
    cin_pad = pad16(Cin)
    if for_dgrad:
        wk = torch.empty((cin_pad, 9 * Cout), dtype=torch.bfloat16, device=w.device)
        _launch("rovr_repack_conv3x3_dgrad", _ptr(w), _ptr(wk), Cout, Cin, cin_pad, _stream())
    else:
        wk = torch.empty((Cout, 9 * cin_pad), dtype=torch.bfloat16, device=w.device)
        _launch("rovr_repack_conv3x3_fprop", _ptr(w), _ptr(wk), Cout, Cin, cin_pad, _stream())
    return wk


def repack_convT2x2(w, for_dgrad=False):
    _f32(w, "convT weight")
    Cin, Cout = w.shape[0], w.shape[1]
    if for_dgrad:
        wk = torch.empty((Cin, 4 * Cout), dtype=torch.bfloat16, device=w.device)
        _launch("rovr_repack_convT2x2_dgrad", _ptr(w), _ptr(wk), Cin, Cout, _stream())
    else:
        wk = torch.empty((4 * Cout, Cin), dtype=torch.bfloat16, device=w.device)
        _launch("rovr_repack_convT2x2_fprop", _ptr(w), _ptr(wk), Cin, Cout, _stream())
    return wk


class _RepackItem(ctypes.Structure):
    _fields_ = [("w", ctypes.c_void_p), ("wk", ctypes.c_void_p), ("kind", ctypes.c_int), ("a", ctypes.c_int),
                ("b", ctypes.c_int), ("c", ctypes.c_int)]


def repack_batch(requests):
    """requests: list of (w, kind, for_dgrad) with kind "conv" (Conv2d 3x3) or "up" (ConvTranspose2d 2x2).
    Returns the bf16 operand copies, all written by ONE kernel launch (rovr_repack_batch)."""
    assert 1 <= len(requests) <= 40
    items = (_RepackItem * len(requests))()
    outs = []
    for it, (w, kind, for_dgrad) in zip(items, requests):
        _f32(w, "weight")
        if kind == "up":
            Cin, Cout = w.shape[0], w.shape[1]
            shape = (Cin, 4 * Cout) if for_dgrad else (4 * Cout, Cin)
            it.kind, it.a, it.b, it.c = (3 if for_dgrad else 2), Cin, Cout, 0
        else:
            Cout, Cin = w.shape[0], w.shape[1]
            cin_pad = pad16(Cin)
            shape = (cin_pad, 9 * Cout) if for_dgrad else (Cout, 9 * cin_pad)
            it.kind, it.a, it.b, it.c = (1 if for_dgrad else 0), Cout, Cin, cin_pad
        wk = torch.empty(shape, dtype=torch.bfloat16, device=w.device)
        it.w, it.wk = _ptr(w), _ptr(wk)
        outs.append(wk)
    _launch("rovr_repack_batch", ctypes.cast(items, ctypes.c_void_p), len(requests), _stream())
    return outs


# ---------------------------------------------------------------------------------------------
# Conv2d 3x3
# ---------------------------------------------------------------------------------------------
def conv3x3_fprop(x, wk, bias, y, relu=True, pooled=None):
    """y = [relu](conv3x3(x) + bias); pooled (optional, [B, H/2, W/2, Cout]) = max_pool2d(y, 2) from the
    same kernel."""
    B, H, W, Cin, x_ld = _act(x, "x")
    By, Hy, Wy, Cout, y_ld = _act(y, "y")
    assert (B, H, W) == (By, Hy, Wy) and wk.shape == (Cout, 9 * Cin), (x.shape, y.shape, wk.shape)
    _f32(bias, "bias")
    if pooled is not None:
        Bp, Hp, Wp, Cp, p_ld = _act(pooled, "pooled")
        assert (Bp, Hp, Wp, Cp) == (B, H // 2, W // 2, Cout)
        _launch("rovr_conv3x3_fprop_pool2", _ptr(x), x_ld, _ptr(wk), _ptr(bias), _ptr(y), y_ld, _ptr(pooled), p_ld,
                B, H, W, Cin, Cout, int(relu), _stream())
        return y
    _launch("rovr_conv3x3_fprop", _ptr(x), x_ld, _ptr(wk), _ptr(bias), _ptr(y), y_ld, B, H, W, Cin,
           Cout, int(relu), _stream())
    return y


def conv3x3_fprop_s2(x, wk, bias, y, relu=True):
    """3x3, padding 1, stride 2: y [B, (H-1)//2+1, (W-1)//2+1, Cout] = [relu](conv(x) + bias)."""
    B, H, W, Cin, x_ld = _act(x, "x")
    By, Hy, Wy, Cout, y_ld = _act(y, "y")
    assert (By, Hy, Wy) == (B, (H - 1) // 2 + 1, (W - 1) // 2 + 1) and wk.shape == (Cout, 9 * Cin), (x.shape, y.shape, wk.shape)
    _f32(bias, "bias")
    _launch("rovr_conv3x3_fprop_s2", _ptr(x), x_ld, _ptr(wk), _ptr(bias), _ptr(y), y_ld, B, H, W, Cin, Cout, int(relu),
            _stream())
    return y


def conv3x3_fprop_s2_f32out(x, wk, bias, y, relu=False):
    """conv3x3_fprop_s2 with an fp32 NHWC output."""
    B, H, W, Cin, x_ld = _act(x, "x")
    By, Hy, Wy, Cout, y_ld = _actf(y, "y")
    assert (By, Hy, Wy) == (B, (H - 1) // 2 + 1, (W - 1) // 2 + 1) and wk.shape == (Cout, 9 * Cin), (x.shape, y.shape, wk.shape)
    _f32(bias, "bias")
    _launch("rovr_conv3x3_fprop_s2_f32out", _ptr(x), x_ld, _ptr(wk), _ptr(bias), _ptr(y), y_ld, B, H, W, Cin, Cout,
            int(relu), _stream())
    return y


def conv3x3_fprop_tail(x, wk, bias, y, w8, b8, target=None):
    """y = relu(conv3x3(x) + bias) (Cout = 64) and, from the same epilogue, out = sigmoid(conv8_1x1(y))
    (NCHW fp32) [+ mean((out - target)^2)]. Returns (out, loss | None)."""
    B, H, W, Cin, x_ld = _act(x, "x")
    By, Hy, Wy, Cout, y_ld = _act(y, "y")
    assert (B, H, W) == (By, Hy, Wy) and Cout == 64 and wk.shape == (Cout, 9 * Cin)
    _f32(bias, "bias")
    w8 = w8.reshape(3, 64)
    _f32(w8, "w8")
    _f32(b8, "b8")
    out = torch.empty((B, 3, H, W), dtype=torch.float32, device=x.device)
    loss = None
    ws, wsn = _vp(0), 0
    if target is not None:
        _f32(target, "target")
        assert target.shape == out.shape
        loss = torch.empty((), dtype=torch.float32, device=x.device)
        wbuf = workspace(N.lib.rovr_conv3x3_fprop_tail_workspace(B, H, W), x.device)
        ws, wsn = _ptr(wbuf), wbuf.numel()
    _launch("rovr_conv3x3_fprop_tail", _ptr(x), x_ld, _ptr(wk), _ptr(bias), _ptr(y), y_ld, _ptr(w8), _ptr(b8),
            _ptr(out), _ptr(target), _ptr(loss), ws, wsn, B, H, W, Cin, Cout, _stream())
    return out, loss


def _colsum_ws(colsum, B, H, W, C, device):
    if colsum is None:
        return _vp(0), 0, _vp(0), 0
    _f32(colsum, "colsum")
    assert 0 < colsum.numel() <= C
    ws = workspace(N.lib.rovr_dgrad_colsum_workspace(B, H, W, C), device)
    return _ptr(colsum), colsum.numel(), _ptr(ws), ws.numel()


def conv3x3_dgrad(dy, wk_d, dx, mask=None, colsum=None, mask_cols=0):
    """dx = conv^T(dy) [* (mask > 0) on the first mask_cols channels (0 = all)]; colsum (fp32, up to Cin
    entries) optionally receives sum_pixels dx of its leading channels."""
    B, H, W, Cout, dy_ld = _act(dy, "dy")
    Bx, Hx, Wx, Cin, dx_ld = _act(dx, "dx")
    assert (B, H, W) == (Bx, Hx, Wx) and wk_d.shape == (Cin, 9 * Cout), (dy.shape, dx.shape, wk_d.shape)
    mask_ld = 0
    if mask is not None:
        Bm, Hm, Wm, Cm, mask_ld = _act(mask, "mask")
        assert (Bm, Hm, Wm, Cm) == (B, H, W, Cin)
    cs, ncs, ws, wsn = _colsum_ws(colsum, B, H, W, Cin, dy.device)
    _launch("rovr_conv3x3_dgrad", _ptr(dy), dy_ld, _ptr(wk_d), _ptr(dx), dx_ld, _ptr(mask), mask_ld, mask_cols,
           B, H, W, Cin, Cout, cs, ncs, ws, wsn, _stream())
    return dx


def conv3x3_wgrad(dy, x, dw):
    B, H, W, Cout, dy_ld = _act(dy, "dy")
    Bx, Hx, Wx, Cin, x_ld = _act(x, "x")
    assert (B, H, W) == (Bx, Hx, Wx)
    _f32(dw, "dw")
    cin_keep = dw.shape[1]
    assert dw.shape == (Cout, cin_keep, 3, 3) and cin_keep <= Cin
    need = N.lib.rovr_conv3x3_wgrad_workspace(B, H, W, Cin, Cout)
    if need == 0:
        raise N.RovrError("conv3x3_wgrad_workspace: " + N.last_error())
    ws = workspace(need, dy.device)
    _launch("rovr_conv3x3_wgrad", _ptr(dy), dy_ld, _ptr(x), x_ld, _ptr(dw), B, H, W, Cin, cin_keep,
           Cout, _ptr(ws), ws.numel(), _stream())
    return dw


# ---------------------------------------------------------------------------------------------
# ConvTranspose2d k2 s2
# ---------------------------------------------------------------------------------------------
def convT2x2_fprop(x, wk, bias, y, relu=True):
    B, H, W, Cin, x_ld = _act(x, "x")
    By, Hy, Wy, Cout, y_ld = _act(y, "y")
    assert (By, Hy, Wy) == (B, 2 * H, 2 * W) and wk.shape == (4 * Cout, Cin), (x.shape, y.shape, wk.shape)
    _f32(bias, "bias")
    _launch("rovr_convT2x2_fprop", _ptr(x), x_ld, _ptr(wk), _ptr(bias), _ptr(y), y_ld, B, H, W, Cin,
           Cout, int(relu), _stream())
    return y


def convT2x2_dgrad(dy, wk_d, dx, mask=None, colsum=None):
    By, Hy, Wy, Cout, dy_ld = _act(dy, "dy")
    B, H, W, Cin, dx_ld = _act(dx, "dx")
    assert (By, Hy, Wy) == (B, 2 * H, 2 * W) and wk_d.shape == (Cin, 4 * Cout)
    assert B == 1 or dy.stride(0) == Hy * Wy * dy_ld
    mask_ld = 0
    if mask is not None:
        Bm, Hm, Wm, Cm, mask_ld = _act(mask, "mask")
        assert (Bm, Hm, Wm, Cm) == (B, H, W, Cin)
    cs, ncs, ws, wsn = _colsum_ws(colsum, B, H, W, Cin, dy.device)
    _launch("rovr_convT2x2_dgrad", _ptr(dy), dy_ld, _ptr(wk_d), _ptr(dx), dx_ld, _ptr(mask), mask_ld,
           B, H, W, Cin, Cout, cs, ncs, ws, wsn, _stream())
    return dx


def convT2x2_wgrad(dy, x, dw):
    By, Hy, Wy, Cout, dy_ld = _act(dy, "dy")
    B, H, W, Cin, x_ld = _act(x, "x")
    assert (By, Hy, Wy) == (B, 2 * H, 2 * W)
    _f32(dw, "dw")
    assert dw.shape == (Cin, Cout, 2, 2)
    need = N.lib.rovr_convT2x2_wgrad_workspace(B, H, W, Cin, Cout)
    if need == 0:
        raise N.RovrError("convT2x2_wgrad_workspace: " + N.last_error())
    ws = workspace(need, dy.device)
    _launch("rovr_convT2x2_wgrad", _ptr(dy), dy_ld, _ptr(x), x_ld, _ptr(dw), B, H, W, Cin, Cout,
           _ptr(ws), ws.numel(), _stream())
    return dw


# ---------------------------------------------------------------------------------------------
# GEMM
# ---------------------------------------------------------------------------------------------
def gemm_bf16(x, wk, bias=None, relu=False, out_dtype=torch.bfloat16, out=None):
    """y[M, N] = x[M, K] @ wk[N, K]^T + bias. x: bf16 2-D (row stride = ld), wk: bf16 [N, K]."""
    assert x.dim() == 2 and x.dtype == torch.bfloat16 and x.stride(1) == 1
    M, K = x.shape
    Nn = wk.shape[0]
    assert wk.shape == (Nn, K) and wk.is_contiguous() and wk.dtype == torch.bfloat16
    if out is None:
        out = torch.empty((M, Nn), dtype=out_dtype, device=x.device)
    assert out.stride(1) == 1
    yb = _ptr(out) if out.dtype == torch.bfloat16 else _vp(0)
    yf = _ptr(out) if out.dtype == torch.float32 else _vp(0)
    _launch("rovr_gemm_bf16", _ptr(x), x.stride(0), _ptr(wk), _ptr(bias), yb, yf, out.stride(0), M,
           Nn, K, int(relu), _stream())
    return out


# ---------------------------------------------------------------------------------------------
# pooling
# ---------------------------------------------------------------------------------------------
def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


def maxpool_fwd(x, y, kernel, stride=None):
    kh, kw = _pair(kernel)
    sh, sw = _pair(stride if stride is not None else kernel)
    B, H, W, C, x_ld = _act(x, "x")
    By, Ho, Wo, Cy, y_ld = _act(y, "y")
    assert (By, Ho, Wo, Cy) == (B, (H - kh) // sh + 1, (W - kw) // sw + 1, C), (x.shape, y.shape)
    _launch("rovr_maxpool_fwd", _ptr(x), x_ld, _ptr(y), y_ld, B, H, W, C, kh, kw, sh, sw, _stream())
    return y


def maxpool_bwd(x, gp, gx, kernel, stride=None, gskip=None, relu_mask=True, colsum=None):
    """gx = [(x > 0) *] (gskip + maxpool_backward(gp)); colsum (fp32 [C]) optionally receives the
    per-channel sums of gx (the bias gradient of the layer that produced x) from the same pass."""
    kh, kw = _pair(kernel)
    sh, sw = _pair(stride if stride is not None else kernel)
    B, H, W, C, x_ld = _act(x, "x")
    _, _, _, _, gp_ld = _act(gp, "gp")
    _, _, _, _, gx_ld = _act(gx, "gx")
    gs_ld = 0
    if gskip is not None:
        _, _, _, _, gs_ld = _act(gskip, "gskip")
    ws, wsn = _vp(0), 0
    if colsum is not None:
        _f32(colsum, "colsum")
        assert colsum.numel() == C
        wbuf = workspace(N.lib.rovr_maxpool_bwd_colsum_workspace(B, H, W, C, kh, kw), x.device)
        ws, wsn = _ptr(wbuf), wbuf.numel()
    _launch("rovr_maxpool_bwd", _ptr(x), x_ld, _ptr(gp), gp_ld, _ptr(gskip), gs_ld, _ptr(gx), gx_ld,
           B, H, W, C, kh, kw, sh, sw, int(relu_mask), _ptr(colsum), ws, wsn, _stream())
    return gx


# ---------------------------------------------------------------------------------------------
# LocalNet tail
# ---------------------------------------------------------------------------------------------
def tail_fwd(y7, w8, b8, target=None):
    """conv8 (1x1, 64->3) + sigmoid. Returns (out NCHW fp32, loss scalar tensor or None)."""
    B, H, W, C, ld = _act(y7, "y7")
    assert C == 64 and ld == 64
    w8 = w8.reshape(3, 64)
    _f32(w8, "w8")
    _f32(b8, "b8")
    out = torch.empty((B, 3, H, W), dtype=torch.float32, device=y7.device)
    loss = None
    ws = workspace(N.lib.rovr_tail_workspace(B, H, W), y7.device)
    if target is not None:
        _f32(target, "target")
        assert target.shape == out.shape
        loss = torch.empty((), dtype=torch.float32, device=y7.device)
    _launch("rovr_tail_fwd", _ptr(y7), _ptr(w8), _ptr(b8), _ptr(out), _ptr(target), _ptr(loss),
           _ptr(ws), ws.numel(), B, H, W, _stream())
    return out, loss


def tail_bwd(y7, w8, out, g7, dw8, db8, gout=None, target=None, mse_scale=0.0, gloss=None, db7=None):
    B, H, W, C, ld = _act(y7, "y7")
    assert C == 64 and ld == 64
    _, _, _, Cg, g_ld = _act(g7, "g7")
    assert Cg == 64 and g_ld == 64
    w8 = w8.reshape(3, 64)
    _f32(w8, "w8")
    _f32(out, "out")
    _f32(gout, "gout")
    _f32(target, "target")
    _f32(gloss, "gloss")
    ws = workspace(N.lib.rovr_tail_workspace(B, H, W), y7.device)
    _launch("rovr_tail_bwd", _ptr(y7), _ptr(w8), _ptr(out), _ptr(gout), _ptr(target),
           ctypes.c_float(mse_scale), _ptr(gloss), _ptr(g7), _ptr(dw8), _ptr(db8), _ptr(db7), _ptr(ws),
           ws.numel(), B, H, W, _stream())


# ---------------------------------------------------------------------------------------------
# bias gradient
# ---------------------------------------------------------------------------------------------
def colsum(g, out):
    B, H, W, C, ld = _act(g, "g")
    _f32(out, "out")
    assert out.numel() == C
    ws = workspace(N.lib.rovr_colsum_workspace(C), g.device)
    _launch("rovr_colsum", _ptr(g), ld, B * H * W, C, _ptr(out), _ptr(ws), ws.numel(), _stream())
    return out


# ---------------------------------------------------------------------------------------------
# linear / 1x1-conv operand packing and tensor-core weight gradient
# ---------------------------------------------------------------------------------------------
def repack_linear(w, transpose=False):
    """fp32 [N, K] (or [N, K, 1, 1]) -> bf16 [pad16(N), pad16(K)], or transposed [pad16(K), pad16(N)]."""
    w = w.reshape(w.shape[0], -1)
    _f32(w, "linear weight")
    Nn, K = w.shape
    n_pad, k_pad = pad16(Nn), pad16(K)
    shape = (k_pad, n_pad) if transpose else (n_pad, k_pad)
    wk = torch.empty(shape, dtype=torch.bfloat16, device=w.device)
    _launch("rovr_repack_linear", _ptr(w), _ptr(wk), Nn, K, n_pad, k_pad, int(transpose), _stream())
    return wk


def gemm_wgrad(dy, x, dw):
    """dw[n_keep, k_keep] (fp32) = dy[M, N]^T @ x[M, K] on tensor cores; dy, x bf16 row-major."""
    assert dy.dim() == 2 and x.dim() == 2 and dy.shape[0] == x.shape[0]
    assert dy.dtype == torch.bfloat16 and x.dtype == torch.bfloat16 and dy.stride(1) == 1 and x.stride(1) == 1
    M, Nn = dy.shape
    K = x.shape[1]
    dw2 = dw.reshape(dw.shape[0], -1)
    _f32(dw2, "dw")
    need = N.lib.rovr_gemm_wgrad_workspace(M, Nn, K)
    if need == 0:
        raise N.RovrError("gemm_wgrad_workspace: " + N.last_error())
    ws = workspace(need, dy.device)
    _launch("rovr_gemm_wgrad", _ptr(dy), dy.stride(0), _ptr(x), x.stride(0), _ptr(dw2), M, Nn, dw2.shape[0],
            K, dw2.shape[1], _ptr(ws), ws.numel(), _stream())
    return dw


# ---------------------------------------------------------------------------------------------
# BatchNorm (train mode) + ReLU, LayerNorm
# ---------------------------------------------------------------------------------------------
def bn_train_fwd(x, y, gamma, beta, c_valid, eps, momentum, running_mean, running_var, nbt, relu=True):
    """Returns (mean, rstd) fp32 [C]; writes y. x / y: NHWC bf16 views with the same shape."""
    B, H, W, C, x_ld = _act(x, "x")
    _, _, _, Cy, y_ld = _act(y, "y")
    assert Cy == C
    mean = torch.empty(C, dtype=torch.float32, device=x.device)
    rstd = torch.empty(C, dtype=torch.float32, device=x.device)
    ws = workspace(N.lib.rovr_bn_workspace(C), x.device)
    _launch("rovr_bn_train_fwd", _ptr(x), x_ld, _ptr(y), y_ld, B * H * W, C, c_valid, _ptr(gamma), _ptr(beta),
            ctypes.c_float(eps), ctypes.c_float(momentum), _ptr(running_mean), _ptr(running_var), _ptr(nbt),
            _ptr(mean), _ptr(rstd), int(relu), _ptr(ws), ws.numel(), _stream())
    return mean, rstd


def bn_train_fwd_frames(x, y, gamma, beta, eps, momentum, running_mean, running_var, nbt, relu=True):
    """Train-mode BatchNorm (+ReLU) with one statistics group per FRAME: the reference's trunk at batch 1 per frame
    (rovr/resnet_extractor.py:42-47). x: the convolution output, NHWC [frames, H, W, C], bf16 (y may then be x) or
    fp32; y: bf16. Running buffers are updated once per frame in frame order. Returns (mean, rstd) fp32 [frames, C]."""
    x_f32 = x.dtype == torch.float32
    B, H, W, C, x_ld = _actf(x, "x") if x_f32 else _act(x, "x")
    By, Hy, Wy, Cy, y_ld = _act(y, "y")
    assert (By, Hy, Wy, Cy) == (B, H, W, C)
    mean = torch.empty((B, C), dtype=torch.float32, device=x.device)
    rstd = torch.empty((B, C), dtype=torch.float32, device=x.device)
    ws = workspace(N.lib.rovr_bn_frames_workspace(C, B, H * W), x.device)
    _launch("rovr_bn_train_fwd_frames", _ptr(x), int(x_f32), x_ld, _ptr(y), y_ld, B, H * W, C, C, _ptr(gamma),
            _ptr(beta), ctypes.c_float(eps), ctypes.c_float(momentum), _ptr(running_mean), _ptr(running_var), _ptr(nbt),
            _ptr(mean), _ptr(rstd), int(relu), _ptr(ws), ws.numel(), _stream())
    return mean, rstd


def bn_eval_fwd(x, y, gamma, beta, c_valid, eps, running_mean, running_var, relu=True):
    """Eval-mode BatchNorm (+ReLU) with the running statistics; returns rstd (fp32 [C]) for backward."""
    B, H, W, C, x_ld = _act(x, "x")
    _, _, _, Cy, y_ld = _act(y, "y")
    assert Cy == C
    rstd = torch.empty(C, dtype=torch.float32, device=x.device)
    _launch("rovr_bn_eval_fwd", _ptr(x), x_ld, _ptr(y), y_ld, B * H * W, C, c_valid, _ptr(gamma), _ptr(beta),
            ctypes.c_float(eps), _ptr(running_mean), _ptr(running_var), _ptr(rstd), int(relu), _stream())
    return rstd


def bn_train_bwd(dy, y, x, dx, gamma, mean, rstd, c_valid, dgamma, dbeta, relu=True, eval_mode=False):
    B, H, W, C, dy_ld = _act(dy, "dy")
    _, _, _, _, y_ld = _act(y, "y")
    _, _, _, _, x_ld = _act(x, "x")
    _, _, _, _, dx_ld = _act(dx, "dx")
    ws = workspace(N.lib.rovr_bn_workspace(C), x.device)
    _launch("rovr_bn_eval_bwd" if eval_mode else "rovr_bn_train_bwd", _ptr(dy), dy_ld, _ptr(y), y_ld, _ptr(x), x_ld,
            _ptr(dx), dx_ld, B * H * W, C,
            c_valid, _ptr(gamma), _ptr(mean), _ptr(rstd), _ptr(dgamma), _ptr(dbeta), int(relu), _ptr(ws),
            ws.numel(), _stream())
    return dx


def layernorm_fwd(x, gamma, beta, eps, want_f32=True, want_bf16=True):
    """x fp32 [..., E] contiguous. Returns (y_f32 | None, y_bf16 | None, mean, rstd)."""
    _f32(x, "x")
    E = x.shape[-1]
    rows = x.numel() // E
    yf = torch.empty_like(x) if want_f32 else None
    yb = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device) if want_bf16 else None
    mean = torch.empty(rows, dtype=torch.float32, device=x.device)
    rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
    _launch("rovr_layernorm_fwd", _ptr(x), rows, E, ctypes.c_float(eps), _ptr(gamma), _ptr(beta), _ptr(yf),
            _ptr(yb), _ptr(mean), _ptr(rstd), _stream())
    return yf, yb, mean, rstd


def layernorm_bwd(g, x, gamma, mean, rstd, dx, accumulate=False, dgamma=None, dbeta=None):
    _f32(g, "g")
    _f32(x, "x")
    E = x.shape[-1]
    rows = x.numel() // E
    ws = workspace(N.lib.rovr_layernorm_workspace(E), x.device)
    _launch("rovr_layernorm_bwd", _ptr(g), _ptr(x), rows, E, _ptr(gamma), _ptr(mean), _ptr(rstd), _ptr(dx),
            int(accumulate), _ptr(dgamma), _ptr(dbeta), _ptr(ws), ws.numel(), _stream())
    return dx


# ---------------------------------------------------------------------------------------------
# fp32 linear (small batch)
# ---------------------------------------------------------------------------------------------
def _rows(t, name):
    if t.dtype != torch.float32 or not t.is_cuda or t.dim() != 2 or t.stride(1) != 1:
        raise ValueError(f"{name}: expected a 2-D CUDA fp32 tensor with unit column stride")
    return t.shape[0], t.shape[1], t.stride(0)


def linear_f32_fwd(x, w, bias, out=None, accumulate=False):
    M, K, x_ld = _rows(x, "x")
    _f32(w, "w")
    Nn = w.shape[0]
    assert w.shape == (Nn, K), (w.shape, x.shape)
    if out is None:
        out = torch.empty((M, Nn), dtype=torch.float32, device=x.device)
    _, _, y_ld = _rows(out, "out")
    _launch("rovr_linear_f32_fwd", _ptr(x), x_ld, _ptr(w), _ptr(bias), _ptr(out), y_ld, M, Nn, K,
            int(accumulate), _stream())
    return out


def linear_f32_dgrad(dy, w, dx=None, accumulate=False):
    M, Nn, dy_ld = _rows(dy, "dy")
    _f32(w, "w")
    K = w.shape[1]
    assert w.shape[0] == Nn
    if dx is None:
        dx = torch.empty((M, K), dtype=torch.float32, device=dy.device)
    assert dx.is_contiguous() and dx.shape == (M, K)
    ws = workspace(N.lib.rovr_linear_f32_dgrad_workspace(M, Nn, K), dy.device)
    _launch("rovr_linear_f32_dgrad", _ptr(dy), dy_ld, _ptr(w), _ptr(dx), M, Nn, K, int(accumulate), _ptr(ws),
            ws.numel(), _stream())
    return dx


def linear_f32_wgrad(dy, x, dw, db=None):
    M, Nn, dy_ld = _rows(dy, "dy")
    _, K, x_ld = _rows(x, "x")
    _f32(dw, "dw")
    assert dw.shape == (Nn, K)
    _launch("rovr_linear_f32_wgrad", _ptr(dy), dy_ld, _ptr(x), x_ld, _ptr(dw), _ptr(db), M, Nn, K, _stream())
    return dw


# ---------------------------------------------------------------------------------------------
# standardisation and policy heads
# ---------------------------------------------------------------------------------------------
def standardize_fwd(x, dim, eps_add):
    """(x - mean) / (std_unbiased + eps_add) along `dim` of a contiguous 2-D fp32 tensor."""
    _f32(x, "x")
    R, C = x.shape
    outer, length, so, si = (R, C, C, 1) if dim == 1 else (C, R, 1, C)
    y = torch.empty_like(x)
    sig = torch.empty(outer, dtype=torch.float32, device=x.device)
    _launch("rovr_standardize_fwd", _ptr(x), _ptr(y), _ptr(sig), outer, length, so, si,
            ctypes.c_float(eps_add), _stream())
    return y, sig


def standardize_bwd(g, y, sig, dim, eps_add):
    _f32(g, "g")
    R, C = y.shape
    outer, length, so, si = (R, C, C, 1) if dim == 1 else (C, R, 1, C)
    dx = torch.empty_like(y)
    _launch("rovr_standardize_bwd", _ptr(g), _ptr(y), _ptr(sig), _ptr(dx), outer, length, so, si,
            ctypes.c_float(eps_add), _stream())
    return dx


def _idx(t, name):
    if t is None:
        return None
    if t.dtype != torch.int64 or not t.is_cuda or not t.is_contiguous():
        raise ValueError(f"{name}: expected a contiguous CUDA int64 tensor")
    return t


def head_mask_std_fwd(logits, target, standardize):
    """In-place scatter of 0 at `target` [b, tk] (or None) on `logits`, then the reference's
    keepdim-less standardisation. Returns (out | None, sig | None)."""
    _f32(logits, "logits")
    b, n = logits.shape
    target = _idx(target, "target")
    tk = 0 if target is None else target.shape[1]
    out = torch.empty_like(logits) if standardize else None
    sig = torch.empty(b, dtype=torch.float32, device=logits.device) if standardize else None
    _launch("rovr_head_mask_std_fwd", _ptr(logits), _ptr(target), tk, _ptr(out), _ptr(sig), b, n,
            int(standardize), _stream())
    return out, sig


def head_mask_std_bwd(g, logits, out, sig, target, standardize):
    _f32(g, "g")
    b, n = logits.shape
    tk = 0 if target is None else target.shape[1]
    dl = torch.empty_like(logits)
    _launch("rovr_head_mask_std_bwd", _ptr(g), _ptr(logits), _ptr(out), _ptr(sig), _ptr(target), tk, _ptr(dl),
            b, n, int(standardize), _stream())
    return dl


def head_gumbel_fwd(logits, expo, tau, mode, action=None):
    """Returns (probs, idx | None, val | None); see rovr_head_gumbel_fwd for the modes."""
    _f32(logits, "logits")
    _f32(expo, "expo")
    b, n = logits.shape
    probs = torch.empty_like(logits)
    idx = val = None
    if mode in (1, 2):
        idx = torch.empty((b,) if mode == 1 else (b, 2), dtype=torch.int64, device=logits.device)
    if mode != 0:
        val = torch.empty(b, dtype=torch.float32, device=logits.device)
    action = _idx(action, "action")
    _launch("rovr_head_gumbel_fwd", _ptr(logits), _ptr(expo), ctypes.c_float(tau), _ptr(probs), b, n, mode,
            _ptr(action), _ptr(idx), _ptr(val), _stream())
    return probs, idx, val


def head_gumbel_bwd(probs, gval, tau, mode, action):
    _f32(gval, "gval")
    b, n = probs.shape
    dl = torch.empty_like(probs)
    _launch("rovr_head_gumbel_bwd", _ptr(probs), _ptr(gval), ctypes.c_float(tau), b, n, mode, _ptr(action),
            _ptr(dl), _stream())
    return dl


# ---------------------------------------------------------------------------------------------
# LSTM pointwise, data movers
# ---------------------------------------------------------------------------------------------
def lstm_pointwise_fwd(gates, c_prev, save_act=True):
    _f32(gates, "gates")
    _f32(c_prev, "c_prev")
    B, Hd = c_prev.shape
    h = torch.empty_like(c_prev)
    c = torch.empty_like(c_prev)
    act = torch.empty_like(gates) if save_act else None
    _launch("rovr_lstm_pointwise_fwd", _ptr(gates), _ptr(c_prev), _ptr(h), _ptr(c), _ptr(act), B, Hd, _stream())
    return h, c, act


def lstm_pointwise_bwd(act, c_prev, c, dh, dc):
    B, Hd = c_prev.shape
    dgates = torch.empty_like(act)
    dc_prev = torch.empty_like(c_prev)
    _launch("rovr_lstm_pointwise_bwd", _ptr(act), _ptr(c_prev), _ptr(c), _ptr(dh), _ptr(dc), _ptr(dgates),
            _ptr(dc_prev), B, Hd, _stream())
    return dgates, dc_prev


def flatten_nhwc(x, C, out, col_off=0):
    """out[b, col_off + c*H*W + p] = x[b, p, c] for c < C (NCHW flatten order)."""
    B, H, W, _, ld = _act(x, "x")
    _, _, out_ld = _rows(out, "out")
    dst = out[:, col_off:]
    _launch("rovr_flatten_nhwc", _ptr(x), ld, _ptr(dst), out_ld, B, H * W, C, _stream())
    return out


def unflatten_nhwc(rows, C, out, col_off=0):
    """out[b, p, c] = rows[b, col_off + c*H*W + p] for c < C, zero for the padding channels."""
    B, H, W, cpad, ld = _act(out, "out")
    _, _, src_ld = _rows(rows, "rows")
    src = rows[:, col_off:]
    _launch("rovr_unflatten_nhwc", _ptr(src), src_ld, _ptr(out), ld, B, H * W, C, cpad, _stream())
    return out


def copy2d_f32(src, dst, scale=1.0, accumulate=False):
    r, c, s_ld = _rows(src, "src")
    r2, c2, d_ld = _rows(dst, "dst")
    assert (r, c) == (r2, c2)
    _launch("rovr_copy2d_f32", _ptr(src), s_ld, _ptr(dst), d_ld, r, c, ctypes.c_float(scale), int(accumulate),
            _stream())
    return dst


# ---------------------------------------------------------------------------------------------
# ResNet-50 frame-feature extractor helpers
# ---------------------------------------------------------------------------------------------
def fold_bn(w, gamma, beta, mean, var, eps):
    """(w * s, beta - mean * s) with s = gamma / sqrt(var + eps); w fp32 [Cout, ...]."""
    _f32(w, "w")
    cout = w.shape[0]
    K = w.numel() // cout
    wf = torch.empty_like(w)
    bf = torch.empty(cout, dtype=torch.float32, device=w.device)
    _launch("rovr_fold_bn", _ptr(w), _ptr(gamma), _ptr(beta), _ptr(mean), _ptr(var), ctypes.c_float(eps),
            _ptr(wf), _ptr(bf), cout, K, _stream())
    return wf, bf


def stem_im2col(x, kpad=160, quantise=True):
    """NCHW fp32 [B,3,H,W] -> ([B*Ho*Wo, kpad] bf16, Ho, Wo) for the 7x7 s2 p3 stem."""
    _f32(x, "x")
    B, C, H, W = x.shape
    assert C == 3
    Ho, Wo = (H + 6 - 7) // 2 + 1, (W + 6 - 7) // 2 + 1
    out = torch.empty((B * Ho * Wo, kpad), dtype=torch.bfloat16, device=x.device)
    _launch("rovr_stem_im2col", _ptr(x), _ptr(out), B, H, W, kpad, int(quantise), _stream())
    return out, Ho, Wo


def maxpool_pad_fwd(x, k, s, pad):
    B, H, W, C, ld = _act(x, "x")
    Ho, Wo = (H + 2 * pad - k) // s + 1, (W + 2 * pad - k) // s + 1
    y = torch.empty((B, Ho, Wo, C), dtype=torch.bfloat16, device=x.device)
    _launch("rovr_maxpool_pad_fwd", _ptr(x), ld, _ptr(y), C, B, H, W, C, k, s, pad, _stream())
    return y


def subsample(x, s):
    B, H, W, C, ld = _act(x, "x")
    Ho, Wo = (H - 1) // s + 1, (W - 1) // s + 1
    y = torch.empty((B, Ho, Wo, C), dtype=torch.bfloat16, device=x.device)
    _launch("rovr_subsample", _ptr(x), ld, _ptr(y), C, B, H, W, C, s, _stream())
    return y


def add_relu(a, b, out=None):
    assert a.shape == b.shape and a.is_contiguous() and b.is_contiguous() and a.dtype == torch.bfloat16
    if out is None:
        out = torch.empty_like(a)
    _launch("rovr_add_relu", _ptr(a), _ptr(b), _ptr(out), a.numel(), _stream())
    return out


def avgpool(x):
    B, H, W, C, ld = _act(x, "x")
    out = torch.empty((B, C), dtype=torch.float32, device=x.device)
    _launch("rovr_avgpool", _ptr(x), ld, _ptr(out), B, H * W, C, _stream())
    return out


def mosaic_paste(feat, mosaic, batch=None, slot=None, slots_per_mosaic=1, tile=16, per_row=5, gather=False):
    """feat [n, ch*tile*tile] fp32 <-> mosaic [nb, ch, side, side] fp32 (in place on `mosaic`, or on
    `feat` when gather=True)."""
    _f32(feat, "feat")
    _f32(mosaic, "mosaic")
    n = feat.shape[0]
    nb, ch, side, _ = mosaic.shape
    assert feat.shape[1] == ch * tile * tile
    _launch("rovr_mosaic_paste", _ptr(feat), _ptr(mosaic), _ptr(_idx(batch, "batch")), _ptr(_idx(slot, "slot")), n,
            slots_per_mosaic, ch, tile, per_row, side, int(gather), _stream())
    return feat if gather else mosaic


def resize_antialias(x, Ho, Wo):
    """ToPILImage -> Resize((Ho, Wo)) -> ToTensor on NCHW fp32 [B,C,H,W] (PIL's 8-bit bilinear)."""
    _f32(x, "x")
    B, C, H, W = x.shape
    out = torch.empty((B, C, Ho, Wo), dtype=torch.float32, device=x.device)
    tmp = torch.empty((B, C, H, Wo), dtype=torch.float32, device=x.device)
    _launch("rovr_resize_antialias", _ptr(x), _ptr(out), _ptr(tmp), B * C, H, W, Ho, Wo, _stream())
    return out


# ---------------------------------------------------------------------------------------------
# attention-block helpers
# ---------------------------------------------------------------------------------------------
def _bf(t, name):
    if t.dtype != torch.bfloat16 or not t.is_cuda:
        raise ValueError(f"{name}: expected a CUDA bf16 tensor")


def gemm_batched(a, b, out):
    """out[i2, i1, m, n] = sum_k a[i2, i1, m, k] * b[i2, i1, n, k].

    a: bf16 view [n2, n1, M, K], b: bf16 view [n2, n1, Nb, K], out: bf16 or fp32 view [n2, n1, M, N]
    with Nb <= N (output columns Nb .. N are zero); arbitrary (8-element-aligned) strides on the
    first three dims, unit stride on the last."""
    _bf(a, "a")
    _bf(b, "b")
    n2, n1, M, K = a.shape
    Nb, Nn = b.shape[2], out.shape[3]
    assert b.shape == (n2, n1, Nb, K) and out.shape == (n2, n1, M, Nn) and Nb <= Nn, (a.shape, b.shape, out.shape)
    assert a.stride(3) == 1 and b.stride(3) == 1 and out.stride(3) == 1
    yb = _ptr(out) if out.dtype == torch.bfloat16 else _vp(0)
    yf = _ptr(out) if out.dtype == torch.float32 else _vp(0)
    _launch("rovr_gemm_batched_bf16", _ptr(a), a.stride(2), a.stride(1), a.stride(0), _ptr(b), b.stride(2),
            b.stride(1), b.stride(0), yb, yf, out.stride(2), out.stride(1), out.stride(0), M, Nn, Nb, K, n1, n2, _stream())
    return out


def transpose_heads(x, r_pad=None):
    """x: bf16 view [n2, n1, R, C] (unit stride on C) -> new contiguous [n2, n1, C, r_pad] with
    out[..., c, r] = x[..., r, c] and zeros for r >= R."""
    _bf(x, "x")
    n2, n1, R, C = x.shape
    r_pad = pad16(R) if r_pad is None else r_pad
    out = torch.empty((n2, n1, C, r_pad), dtype=torch.bfloat16, device=x.device)
    assert x.stride(3) == 1
    _launch("rovr_transpose_bf16", _ptr(x), x.stride(2), x.stride(1), x.stride(0), _ptr(out), out.stride(2),
            out.stride(1), out.stride(0), R, C, r_pad, n1, n2, _stream())
    return out


def softmax_fwd(s, T, scale):
    """s: fp32 contiguous [..., t_pad] -> bf16 probabilities, zero for t >= T."""
    _f32(s, "s")
    t_pad = s.shape[-1]
    p = torch.empty(s.shape, dtype=torch.bfloat16, device=s.device)
    _launch("rovr_softmax_fwd", _ptr(s), _ptr(p), s.numel() // t_pad, T, t_pad, ctypes.c_float(scale), _stream())
    return p


def softmax_bwd(dp, p, T, scale):
    _f32(dp, "dp")
    t_pad = p.shape[-1]
    ds = torch.empty_like(p)
    _launch("rovr_softmax_bwd", _ptr(dp), _ptr(p), _ptr(ds), p.numel() // t_pad, T, t_pad, ctypes.c_float(scale),
            _stream())
    return ds


def gelu_fwd(h):
    _bf(h, "h")
    assert h.is_contiguous()
    a = torch.empty_like(h)
    _launch("rovr_gelu_fwd", _ptr(h), _ptr(a), h.numel(), _stream())
    return a


def gelu_bwd(da, h):
    assert da.is_contiguous() and h.is_contiguous()
    dh = torch.empty_like(h)
    _launch("rovr_gelu_bwd", _ptr(da), _ptr(h), _ptr(dh), h.numel(), _stream())
    return dh


def cast_bf16(x):
    _f32(x, "x")
    out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    _launch("rovr_cast_f32_bf16", _ptr(x), _ptr(out), x.numel(), _stream())
    return out


def add_f32(a, b):
    """a + b for two dense fp32 tensors of the same shape (one pass)."""
    _f32(a, "a")
    _f32(b, "b")
    assert a.shape == b.shape and a.numel() % 4 == 0
    out = torch.empty_like(a)
    _launch("rovr_add_f32", _ptr(a), _ptr(b), _ptr(out), a.numel(), _stream())
    return out


def colsum_rows(g, out):
    """out[c] = sum_m g[m, c] for a bf16 [M, C] matrix of any even width (one kernel + one reduction)."""
    _bf(g, "g")
    M, C = g.shape
    assert out.numel() == C and g.stride(1) == 1
    _f32(out, "out")
    ws = workspace(N.lib.rovr_colsum_rows_workspace(C), g.device)
    _launch("rovr_colsum_rows", _ptr(g), g.stride(0), M, C, _ptr(out), _ptr(ws), ws.numel(), _stream())
    return out


def linear_wgrad(dy, x, dw):
    """dw[N, K] (fp32) = dy[M, N]^T @ x[M, K] for large M: both operands are transposed once (zero-padded
    to a multiple of 16 rows) and the product runs as a regular tcgen05 GEMM with full 128 x 256 tiles
    and an fp32 epilogue straight into dw — no split-K partials. For small M use gemm_wgrad."""
    _bf(dy, "dy")
    _bf(x, "x")
    M, Nn = dy.shape
    K = x.shape[1]
    assert x.shape[0] == M and dw.shape == (Nn, K) and dw.dtype == torch.float32 and dw.is_contiguous()
    m_pad = pad16(M)
    dyT = transpose_heads(dy.unflatten(0, (1, 1, M)), m_pad)[0, 0]      # [N, m_pad]
    xT = transpose_heads(x.unflatten(0, (1, 1, M)), m_pad)[0, 0]        # [K, m_pad]
    gemm_bf16(dyT, xT, None, out=dw)
    return dw


def posenc_add(x, w1, b1, n1, w2=None, b2=None):
    _f32(x, "x")
    B, P, D = x.shape
    out = torch.empty_like(x)
    _launch("rovr_posenc_add", _ptr(x), _ptr(out), _ptr(w1), _ptr(b1), _ptr(w2), _ptr(b2), B, P, D, n1, _stream())
    return out


def posenc_grad(g, n1, which):
    _f32(g, "g")
    B, P, D = g.shape
    gw = torch.empty(D, dtype=torch.float32, device=g.device)
    gb = torch.empty(D, dtype=torch.float32, device=g.device)
    _launch("rovr_posenc_grad", _ptr(g), B, P, D, n1, which, _ptr(gw), _ptr(gb), _stream())
    return gw, gb


# ---------------------------------------------------------------------------------------------
# emulated-fp32 policy trunks (csrc/fp32x.cuh): fp32 NHWC activations, split-bf16 stacked operands
# ---------------------------------------------------------------------------------------------
def _actf(t, name="activation", align=4):
    """Validate an NHWC fp32 view (pixel stride a multiple of `align` floats) and return (B, H, W, C, ld)."""
    if t.dtype != torch.float32 or not t.is_cuda or t.dim() != 4:
        raise ValueError(f"{name}: expected a 4-D CUDA fp32 NHWC tensor, got {t.dtype} {tuple(t.shape)}")
    B, H, W, C = t.shape
    ld = t.stride(2)
    if t.stride(3) != 1 or t.stride(1) != W * ld or (B > 1 and t.stride(0) != H * W * ld) or ld % align:
        raise ValueError(f"{name}: not a dense-pixel NHWC view, strides={t.stride()}")
    return B, H, W, C, ld


def pad8(c):
    return (c + 7) // 8 * 8


def split_stack(src, nterms, cb, layout="nhwc", pool=None, out=None, src2=None, relu_mask=None):
    """Stacked bf16 pieces of an fp32 tensor: out[b, y, x, t*cb + c] = piece_t(maxpool(src)[b, y, x, c]),
    zero for c >= C. src: NHWC fp32 view (layout "nhwc") or NCHW fp32 tensor ("nchw"; src2 = a second
    NCHW tensor of the same shape concatenated behind it along C). pool: None or (kh, kw, sh, sw).
    nterms 6 -> forward operand [h m l h m h], 3 -> gradient operand [h m h]."""
    c_split = 0
    if layout == "nhwc":
        B, H, W, C, ld = _actf(src, "split source", align=1)     # scalar loads: any pixel stride
        sb, sy, sx, sc = H * W * ld, W * ld, ld, 1
        assert src2 is None
    else:
        _f32(src, "split source")
        B, C, H, W = src.shape
        sb, sy, sx, sc = C * H * W, W, 1, H * W
        if src2 is not None:
            _f32(src2, "split source 2")
            assert src2.shape == src.shape
            c_split, C = C, 2 * C
    kh, kw, sh, sw = pool if pool is not None else (1, 1, 1, 1)
    Ho, Wo = (H - kh) // sh + 1, (W - kw) // sw + 1
    assert cb >= C and cb % 2 == 0
    if out is None:
        out = torch.empty((B, Ho, Wo, nterms * cb), dtype=torch.bfloat16, device=src.device)
    if relu_mask is not None:       # same view geometry as src: value taken as 0 where mask <= 0
        assert layout == "nhwc" and pool is None and relu_mask.shape == src.shape and relu_mask.stride() == src.stride()
        assert relu_mask.dtype == torch.float32
    _launch("rovr_split_stack", _ptr(src), _ptr(src2), c_split, _ptr(relu_mask), sb, sy, sx, sc, B, H, W, C, kh, kw, sh, sw, _ptr(out), out.stride(2), cb, 0,
            cb, nterms, _stream())
    return out


def split_weights(w, stack_dim, nterms, cb):
    """fp32 weight [d0, d1, ...] -> fp32 [.., nterms*cb, ..] stacked pieces along dim 0 or 1 (zero padded)."""
    _f32(w, "weight")
    d0, d1 = w.shape[0], w.shape[1]
    inner = w.numel() // (d0 * d1)
    shape = list(w.shape)
    shape[stack_dim] = nterms * cb
    out = torch.empty(shape, dtype=torch.float32, device=w.device)
    _launch("rovr_split_weights", _ptr(w), _ptr(out), d0, d1, inner, stack_dim, nterms, cb, _stream())
    return out


def blocksum4(dwp, dw, cb0, cb1):
    """dw[i, j, ...] = sum of the four [cb0, cb1] blocks of dwp [2*cb0, 2*cb1, ...]."""
    _f32(dwp, "dwp")
    _f32(dw, "dw")
    d0, d1 = dw.shape[0], dw.shape[1]
    inner = dw.numel() // (d0 * d1)
    assert dwp.shape[0] == 2 * cb0 and dwp.shape[1] == 2 * cb1 and dwp.numel() == 4 * cb0 * cb1 * inner
    _launch("rovr_blocksum4", _ptr(dwp), _ptr(dw), d0, d1, inner, cb0, cb1, _stream())
    return dw


def conv3x3_f32out(x, wk, bias, y, relu=False):
    """y (fp32 NHWC view) = conv3x3(x bf16 stacked operand) + bias; also the data gradient when wk is the
    dgrad-packed operand (then x = stacked dy and y = dx)."""
    B, H, W, Cin, x_ld = _act(x, "x")
    By, Hy, Wy, Cout, y_ld = _actf(y, "y")
    assert (B, H, W) == (By, Hy, Wy) and wk.shape == (Cout, 9 * Cin), (x.shape, y.shape, wk.shape)
    _f32(bias, "bias")
    _launch("rovr_conv3x3_f32out", _ptr(x), x_ld, _ptr(wk), _ptr(bias), _ptr(y), y_ld, B, H, W, Cin, Cout, int(relu),
            _stream())
    return y


def convT2x2_fprop_f32out(x, wk, bias, y, relu=False):
    B, H, W, Cin, x_ld = _act(x, "x")
    By, Hy, Wy, Cout, y_ld = _actf(y, "y")
    assert (By, Hy, Wy) == (B, 2 * H, 2 * W) and wk.shape == (4 * Cout, Cin), (x.shape, y.shape, wk.shape)
    _f32(bias, "bias")
    _launch("rovr_convT2x2_fprop_f32out", _ptr(x), x_ld, _ptr(wk), _ptr(bias), _ptr(y), y_ld, B, H, W, Cin, Cout,
            int(relu), _stream())
    return y


def convT2x2_dgrad_f32out(dy, wk_d, dx):
    By, Hy, Wy, Cout, dy_ld = _act(dy, "dy")
    B, H, W, Cin, dx_ld = _actf(dx, "dx")
    assert (By, Hy, Wy) == (B, 2 * H, 2 * W) and wk_d.shape == (Cin, 4 * Cout), (dy.shape, dx.shape, wk_d.shape)
    _launch("rovr_convT2x2_dgrad_f32out", _ptr(dy), dy_ld, _ptr(wk_d), _ptr(dx), dx_ld, B, H, W, Cin, Cout, _stream())
    return dx


def bn_f32_train_fwd(x, y, gamma, beta, c_valid, eps, momentum, running_mean, running_var, nbt, relu=True):
    B, H, W, C, x_ld = _actf(x, "x")
    _, _, _, Cy, y_ld = _actf(y, "y")
    assert Cy == C
    mean = torch.empty(C, dtype=torch.float32, device=x.device)
    rstd = torch.empty(C, dtype=torch.float32, device=x.device)
    ws = workspace(N.lib.rovr_bn_workspace(C), x.device)
    _launch("rovr_bn_f32_train_fwd", _ptr(x), x_ld, _ptr(y), y_ld, B * H * W, C, c_valid, _ptr(gamma), _ptr(beta),
            ctypes.c_float(eps), ctypes.c_float(momentum), _ptr(running_mean), _ptr(running_var), _ptr(nbt),
            _ptr(mean), _ptr(rstd), int(relu), _ptr(ws), ws.numel(), _stream())
    return mean, rstd


def bn_f32_eval_fwd(x, y, gamma, beta, c_valid, eps, running_mean, running_var, relu=True):
    B, H, W, C, x_ld = _actf(x, "x")
    _, _, _, Cy, y_ld = _actf(y, "y")
    assert Cy == C
    rstd = torch.empty(C, dtype=torch.float32, device=x.device)
    _launch("rovr_bn_f32_eval_fwd", _ptr(x), x_ld, _ptr(y), y_ld, B * H * W, C, c_valid, _ptr(gamma), _ptr(beta),
            ctypes.c_float(eps), _ptr(running_mean), _ptr(running_var), _ptr(rstd), int(relu), _stream())
    return rstd


def bn_f32_bwd(dy, y, x, dx, gamma, mean, rstd, c_valid, dgamma, dbeta, relu=True, eval_mode=False):
    B, H, W, C, dy_ld = _actf(dy, "dy")
    _, _, _, _, y_ld = _actf(y, "y")
    _, _, _, _, x_ld = _actf(x, "x")
    _, _, _, _, dx_ld = _actf(dx, "dx")
    ws = workspace(N.lib.rovr_bn_workspace(C), x.device)
    _launch("rovr_bn_f32_bwd", _ptr(dy), dy_ld, _ptr(y), y_ld, _ptr(x), x_ld, _ptr(dx), dx_ld, B * H * W, C, c_valid,
            _ptr(gamma), _ptr(mean), _ptr(rstd), _ptr(dgamma), _ptr(dbeta), int(relu), int(eval_mode), _ptr(ws),
            ws.numel(), _stream())
    return dx


def colsum_f32(g, out):
    B, H, W, C, ld = _actf(g, "g")
    _f32(out, "out")
    assert out.numel() == C
    ws = workspace(N.lib.rovr_bn_workspace(C), g.device)
    _launch("rovr_colsum_f32", _ptr(g), ld, B * H * W, C, _ptr(out), _ptr(ws), ws.numel(), _stream())
    return out


def maxpool_f32_fwd(x, kernel, stride=None, out=None):
    kh, kw = _pair(kernel)
    sh, sw = _pair(stride if stride is not None else kernel)
    B, H, W, C, x_ld = _actf(x, "x")
    Ho, Wo = (H - kh) // sh + 1, (W - kw) // sw + 1
    if out is None:
        out = torch.empty((B, Ho, Wo, C), dtype=torch.float32, device=x.device)
    _, _, _, _, y_ld = _actf(out, "y")
    _launch("rovr_maxpool_f32_fwd", _ptr(x), x_ld, _ptr(out), y_ld, B, H, W, C, kh, kw, sh, sw, _stream())
    return out


def maxpool_f32_bwd(x, gp, gx, kernel, stride=None, gskip=None):
    """gx = [gskip +] maxpool_backward(gp) with the arg-max recomputed from the fp32 activation x."""
    kh, kw = _pair(kernel)
    sh, sw = _pair(stride if stride is not None else kernel)
    B, H, W, C, x_ld = _actf(x, "x")
    _, _, _, _, gp_ld = _actf(gp, "gp")
    _, _, _, _, gx_ld = _actf(gx, "gx")
    gs_ld = 0
    if gskip is not None:
        _, _, _, _, gs_ld = _actf(gskip, "gskip")
    _launch("rovr_maxpool_f32_bwd", _ptr(x), x_ld, _ptr(gp), gp_ld, _ptr(gskip), gs_ld, _ptr(gx), gx_ld, B, H, W, C,
            kh, kw, sh, sw, _stream())
    return gx


def flatten_f32(x, C, out, col_off=0):
    """out[b, col_off + c*H*W + p] = x[b, p, c] for c < C (nn.Flatten of the NCHW view)."""
    B, H, W, _, ld = _actf(x, "x")
    _, _, out_ld = _rows(out, "out")
    _launch("rovr_flatten_f32", _ptr(x), ld, _ptr(out[:, col_off:]), out_ld, B, H * W, C, _stream())
    return out


def unflatten_f32(rows, C, out, col_off=0):
    B, H, W, cpad, ld = _actf(out, "out")
    _, _, src_ld = _rows(rows, "rows")
    _launch("rovr_unflatten_f32", _ptr(rows[:, col_off:]), src_ld, _ptr(out), ld, B, H * W, C, cpad, _stream())
    return out


# ---------------------------------------------------------------------------------------------
# LPIPS(net='vgg') head (csrc/lpips.cuh)
# ---------------------------------------------------------------------------------------------
def _f3(vals):
    return (ctypes.c_float * 3)(*[float(v) for v in vals])


def lpips_pack(in0, in1, shift, scale, normalize):
    """ScalingLayer of both images -> one NHWC bf16 batch [2N, H, W, 16] (in0 first)."""
    _f32(in0, "in0")
    _f32(in1, "in1")
    N, C, H, W = in0.shape
    assert C == 3 and in1.shape == in0.shape
    out = torch.empty((2 * N, H, W, 16), dtype=torch.bfloat16, device=in0.device)
    _launch("rovr_lpips_pack", _ptr(in0), _ptr(in1), _ptr(out), N, H, W, _f3(shift), _f3(scale), int(normalize), _stream())
    return out


def lpips_head(feats, lin_w, want_grad):
    """feats [2N, h, w, C] bf16 -> (partial [N, nblocks] fp32, grad [N, h, w, C] bf16 | None)."""
    B2, h, w, C, ld = _act(feats, "feats")
    assert ld == C and B2 % 2 == 0 and feats.is_contiguous()
    N = B2 // 2
    _f32(lin_w, "lin_w")
    assert lin_w.numel() == C
    nb = N_lib_head_blocks(N, h * w)
    partial = torch.empty((N, nb), dtype=torch.float32, device=feats.device)
    grad = torch.empty((N, h, w, C), dtype=torch.bfloat16, device=feats.device) if want_grad else None
    _launch("rovr_lpips_head", _ptr(feats), N, h * w, C, _ptr(lin_w), _ptr(grad), _ptr(partial), nb, _stream())
    return partial, grad


def N_lib_head_blocks(N_img, hw):
    nb = N.lib.rovr_lpips_head_blocks(N_img, hw)
    if nb <= 0:
        raise N.RovrError("lpips_head_blocks: " + N.last_error())
    return nb


def lpips_finalize(partials, hws, n_img):
    """partials: list of [N, nblocks] fp32 tensors (one per tap), hws: pixels per tap -> val [N] fp32."""
    k = len(partials)
    ptrs = (ctypes.c_void_p * k)(*[p.data_ptr() for p in partials])
    nbs = (ctypes.c_int * k)(*[p.shape[1] for p in partials])
    hw = (ctypes.c_longlong * k)(*[int(v) for v in hws])
    val = torch.empty(n_img, dtype=torch.float32, device=partials[0].device)
    for p in partials:
        _ptr(p)
    _launch("rovr_lpips_finalize", ctypes.cast(ptrs, ctypes.c_void_p), ctypes.cast(nbs, ctypes.c_void_p),
            ctypes.cast(hw, ctypes.c_void_p), k, n_img, _ptr(val), _stream())
    return val


def lpips_unpack_grad(gx16, gval, shift, scale, normalize):
    """gradient w.r.t. the packed input [N, H, W, 16] bf16 -> d/d in0 [N, 3, H, W] fp32, times gval[n]."""
    Nn, H, W, C, ld = _act(gx16, "gx16")
    assert C == 16 and ld == 16
    _f32(gval, "gval")
    assert gval.numel() == Nn
    out = torch.empty((Nn, 3, H, W), dtype=torch.float32, device=gx16.device)
    _launch("rovr_lpips_unpack_grad", _ptr(gx16), _ptr(gval), _ptr(out), Nn, H, W, _f3(shift), _f3(scale),
            int(normalize), _stream())
    return out


# ---------------------------------------------------------------------------------------------
# dropout with a recomputable counter-based mask
# ---------------------------------------------------------------------------------------------
def dropout(x, p, state, site, out=None):
    """out = x * keep / (1 - p); keep = f(state {seed, counter} on the device, site, element index).
    x: contiguous bf16 or fp32. The same (state, site) gives the same mask (backward recomputes it)."""
    assert x.is_contiguous() and x.dtype in (torch.bfloat16, torch.float32)
    assert state.dtype == torch.int64 and state.numel() == 2 and state.is_cuda
    if out is None:
        out = torch.empty_like(x)
    _launch("rovr_dropout", _ptr(x), _ptr(out), x.numel(), int(x.dtype == torch.float32), ctypes.c_float(p), _ptr(state),
            int(site), _stream())
    return out


def dropout_advance(state):
    _launch("rovr_dropout_advance", _ptr(state), _stream())


def lpips_pack_f32(in0, in1, shift, scale, normalize):
    """fp32 variant: [2N, H, W, 4] fp32 (3 valid channels)."""
    _f32(in0, "in0")
    _f32(in1, "in1")
    N_, C, H, W = in0.shape
    assert C == 3 and in1.shape == in0.shape
    out = torch.empty((2 * N_, H, W, 4), dtype=torch.float32, device=in0.device)
    _launch("rovr_lpips_pack_f32", _ptr(in0), _ptr(in1), _ptr(out), N_, H, W, _f3(shift), _f3(scale), int(normalize), _stream())
    return out


def lpips_head_f32(feats, lin_w, want_grad):
    B2, h, w, C, ld = _actf(feats, "feats")
    assert ld == C and B2 % 2 == 0 and feats.is_contiguous()
    N_ = B2 // 2
    nb = N_lib_head_blocks(N_, h * w)
    partial = torch.empty((N_, nb), dtype=torch.float32, device=feats.device)
    grad = torch.empty((N_, h, w, C), dtype=torch.float32, device=feats.device) if want_grad else None
    _launch("rovr_lpips_head_f32", _ptr(feats), N_, h * w, C, _ptr(lin_w), _ptr(grad), _ptr(partial), nb, _stream())
    return partial, grad


def lpips_unpack_grad_f32(gx, gval, shift, scale, normalize):
    Nn, H, W, C, ld = _actf(gx, "gx")
    _f32(gval, "gval")
    out = torch.empty((Nn, 3, H, W), dtype=torch.float32, device=gx.device)
    _launch("rovr_lpips_unpack_grad_f32", _ptr(gx), ld, _ptr(gval), _ptr(out), Nn, H, W, _f3(shift), _f3(scale),
            int(normalize), _stream())
    return out
