"""Drop-in for the reference `common_layers` module (rovr/common_layers.py:7-118):
`ImagePositionalEncoding`, `ContextPositionalEncoding`, `SelfAttentionBlock`,
`CrossAttentionBlock`, `FeedForwardBlock`, `EncoderBlock`, `DecoderBlock` — same class names,
constructor arguments, `forward` signatures, sub-module names and state_dict keys
(`attention.in_proj_weight`, `attention.out_proj.weight`, `layer_norm.weight`, `fc1.weight`, ...).

Each block is one torch.autograd.Function over the B200 kernels:
  LayerNorm                       warp-per-row kernel (fp32 in, fp32 + bf16 out)
  packed QKV / output / FFN GEMMs tcgen05 GEMM (rovr_gemm_bf16), bias (+fp32 output) in the epilogue
  QK^T and PV per (batch, head)   tcgen05 batched GEMM fed by rank-5 TMA maps straight out of the
                                  packed [tokens, 3E] projection buffer (no head split copies)
  softmax over key tokens         warp-per-row kernel (scale 1/sqrt(d) folded in), exact GELU kernel
  weight gradients                tcgen05 split-K wgrad engine; bias gradients = column sums
and the backward pass is scheduled by hand. The residual of the attention blocks is taken on the
NORMALISED input, as the reference does (x = LN(x); x = x + MHA(x), :62-63).

Dropout (training mode, p > 0): nn.MultiheadAttention drops attention probabilities (:58,70) and
FeedForwardBlock drops the GELU output (:87,91). Masks come from a counter-based hash of
(seed, step counter, call site, element index) kept in a device int64[2] per block: nothing is
stored — backward recomputes the mask — and a CUDA-graph replay draws a fresh mask because the
counter increment is captured with the step. The seed is drawn from torch's CPU generator at first
use (torch.manual_seed makes runs reproducible); the masks are NOT torch's own Philox stream, so
parity tests feed the module's mask (`ops.dropout` on ones) to the oracle.
"""
import math

import torch
import torch.nn as nn

import ops
from _blocks import PackedWeights

BF = torch.bfloat16


def _check_dims(E, heads):
    if E % 16 or (E // heads) % 16 or E % heads:
        raise ValueError(f"B200 attention path needs hidden_dim and head_dim to be multiples of 16 (E={E}, heads={heads})")


def _need_cuda(x, who):
    if not x.is_cuda:
        raise RuntimeError(f"{who} (B200) needs CUDA tensors: there is no CPU path")


def _heads_view(buf, b, n_tok, heads, d, col0):
    """[b*n_tok, W] projection buffer -> view [b, heads, n_tok, d] of columns col0 .. col0 + heads*d."""
    W = buf.shape[1]
    return buf.as_strided((b, heads, n_tok, d), (n_tok * W, d, W, 1), buf.storage_offset() + col0)


def _attention_core_fwd(q_buf, q_col, kv_buf, k_col, v_col, b, S, T, heads, d, drop=None):
    """softmax(Q K^T / sqrt(d)) V per (batch, head). Returns (O [b*S, E] bf16, P [b, h, S, t_pad] bf16).
    drop = (p, state snapshot, site): attention-probability dropout of nn.MultiheadAttention."""
    dev = q_buf.device
    E = heads * d
    t_pad = ops.pad16(T)
    Q = _heads_view(q_buf, b, S, heads, d, q_col)
    K = _heads_view(kv_buf, b, T, heads, d, k_col)
    V = _heads_view(kv_buf, b, T, heads, d, v_col)
    scores = torch.empty((b, heads, S, t_pad), dtype=torch.float32, device=dev)
    ops.gemm_batched(Q, K, scores)                                   # rows of K beyond T are TMA zero fill
    P = ops.softmax_fwd(scores, T, 1.0 / math.sqrt(d))
    Vt = ops.transpose_heads(V, t_pad)                               # [b, h, d, t_pad]
    O = torch.empty((b * S, E), dtype=BF, device=dev)
    Pd = P if drop is None else ops.dropout(P, *drop)
    ops.gemm_batched(Pd, Vt, _heads_view(O, b, S, heads, d, 0))
    return O, P


def _attention_core_bwd(dO, P, q_buf, q_col, kv_buf, k_col, v_col, dq_buf, dk_buf_view, dv_buf_view, b, S, T, heads, d,
                        drop=None):
    """Gradients of the core: writes dQ / dK / dV (views [b, h, tokens, d] of the projection-gradient buffers)."""
    dev = dO.device
    t_pad, s_pad = ops.pad16(T), ops.pad16(S)
    Q = _heads_view(q_buf, b, S, heads, d, q_col)
    K = _heads_view(kv_buf, b, T, heads, d, k_col)
    V = _heads_view(kv_buf, b, T, heads, d, v_col)
    dOh = _heads_view(dO, b, S, heads, d, 0)
    # dV = P^T dO
    Pd = P if drop is None else ops.dropout(P, *drop)                      # the mask is recomputed, not stored
    Pt = ops.transpose_heads(Pd[..., :T] if t_pad != T else Pd, s_pad)     # [b, h, T, s_pad]
    dOt = ops.transpose_heads(dOh, s_pad)                                  # [b, h, d, s_pad]
    ops.gemm_batched(Pt, dOt, dv_buf_view)
    # dP = dO V^T ; dS = softmax'(dP)
    dP = torch.empty((b, heads, S, t_pad), dtype=torch.float32, device=dev)
    ops.gemm_batched(dOh, V, dP)
    if drop is not None:
        ops.dropout(dP, *drop, out=dP)                                     # d/dP of P * keep / (1 - p)
    dS = ops.softmax_bwd(dP, P, T, 1.0 / math.sqrt(d))
    # dQ = dS K ; dK = dS^T Q
    Kt = ops.transpose_heads(K, t_pad)                                     # [b, h, d, t_pad]
    ops.gemm_batched(dS, Kt, dq_buf)
    dSt = ops.transpose_heads(dS[..., :T] if t_pad != T else dS, s_pad)    # [b, h, T, s_pad]
    Qt = ops.transpose_heads(Q, s_pad)                                     # [b, h, d, s_pad]
    ops.gemm_batched(dSt, Qt, dk_buf_view)


class _Packed:
    """bf16 [N, K] and transposed [K, N] operand copies of nn.Linear-style weights."""

    def __init__(self):
        self.cache = PackedWeights()

    def fwd(self, key, w):
        return self.cache.get((key, "f"), w, lambda t: ops.repack_linear(t, False))

    def bwd(self, key, w):
        return self.cache.get((key, "d"), w, lambda t: ops.repack_linear(t, True))


def _wgrad(gy_bf, x_bf, gw):
    """dW = gy^T x: a regular GEMM over transposed operands when there are many tokens (full MMA tiles,
    no split-K partial traffic), the split-K wgrad engine otherwise."""
    if gy_bf.shape[0] >= 512 and gw.is_contiguous():
        ops.linear_wgrad(gy_bf, x_bf, gw)
    else:
        ops.gemm_wgrad(gy_bf, x_bf, gw)


def _linear_grads(gy_bf, x_bf, w, want_bias=True):
    gw = torch.empty_like(w)
    _wgrad(gy_bf, x_bf, gw)
    gb = None
    if want_bias:
        gb = torch.empty(w.shape[0], dtype=torch.float32, device=w.device)
        ops.colsum_rows(gy_bf, gb)
    return gw, gb


class _AttentionFn(torch.autograd.Function):
    """x (and optionally enc) -> LN(x) + MHA(LN(x), kv, kv)   (rovr/common_layers.py:61-64,73-78)."""

    @staticmethod
    def forward(ctx, blk, x, enc, ln_w, ln_b, lne_w, lne_b, w_in, b_in, w_out, b_out):
        heads = blk.attention.num_heads
        b, S, E = x.shape
        d = E // heads
        cross = enc is not None
        pk = blk._packed
        xn32, xn16, mean, rstd = ops.layernorm_fwd(x.contiguous(), ln_w, ln_b, blk.layer_norm.eps)
        x2 = xn16.view(b * S, E)
        win = pk.fwd("in", w_in)
        if cross:
            T = enc.shape[1]
            _, en16, emean, erstd = ops.layernorm_fwd(enc.contiguous(), lne_w, lne_b, blk.layer_norm_encoder_output.eps,
                                                      want_f32=False)
            e2 = en16.view(b * T, E)
            q_buf = ops.gemm_bf16(x2, win[:E], b_in[:E].contiguous())
            kv_buf = ops.gemm_bf16(e2, win[E:], b_in[E:].contiguous())
            q_col, k_col, v_col = 0, 0, E
        else:
            T = S
            e2 = emean = erstd = None
            q_buf = kv_buf = ops.gemm_bf16(x2, win, b_in)
            q_col, k_col, v_col = 0, E, 2 * E
        drop = _draw_dropout(blk, blk.attention.dropout, 0)
        O, P = _attention_core_fwd(q_buf, q_col, kv_buf, k_col, v_col, b, S, T, heads, d, drop)
        y = ops.gemm_bf16(O, pk.fwd("out", w_out), b_out, out_dtype=torch.float32)
        ops.copy2d_f32(xn32.view(b * S, E), y, accumulate=True)            # residual on the normalised x
        ctx.blk, ctx.cross, ctx.dims = blk, cross, (b, S, T, E, heads, d)
        ctx.cols = (q_col, k_col, v_col)
        ctx.drop = drop
        ctx.save_for_backward(x, enc, ln_w, lne_w, w_in, w_out, mean, rstd, emean, erstd, x2, e2, q_buf, kv_buf, O, P)
        return y.view(b, S, E)

    @staticmethod
    def backward(ctx, g):
        (x, enc, ln_w, lne_w, w_in, w_out, mean, rstd, emean, erstd, x2, e2, q_buf, kv_buf, O, P) = ctx.saved_tensors
        blk, cross = ctx.blk, ctx.cross
        b, S, T, E, heads, d = ctx.dims
        q_col, k_col, v_col = ctx.cols
        pk = blk._packed
        dev = g.device
        g = g.contiguous().float()
        g16 = ops.cast_bf16(g.view(b * S, E))
        gw_out, gb_out = _linear_grads(g16, O, w_out)
        dO = ops.gemm_bf16(g16, pk.bwd("out", w_out))
        if cross:
            dq_buf = torch.empty((b * S, E), dtype=BF, device=dev)
            dkv_buf = torch.empty((b * T, 2 * E), dtype=BF, device=dev)
            dq_v = _heads_view(dq_buf, b, S, heads, d, 0)
            dk_v = _heads_view(dkv_buf, b, T, heads, d, 0)
            dv_v = _heads_view(dkv_buf, b, T, heads, d, E)
        else:
            dq_buf = dkv_buf = torch.empty((b * S, 3 * E), dtype=BF, device=dev)
            dq_v = _heads_view(dq_buf, b, S, heads, d, 0)
            dk_v = _heads_view(dq_buf, b, S, heads, d, E)
            dv_v = _heads_view(dq_buf, b, S, heads, d, 2 * E)
        _attention_core_bwd(dO, P, q_buf, q_col, kv_buf, k_col, v_col, dq_v, dk_v, dv_v, b, S, T, heads, d, ctx.drop)
        gw_in = torch.empty_like(w_in)
        gb_in = torch.empty(3 * E, dtype=torch.float32, device=dev)
        genc = glne_w = glne_b = None
        if cross:
            _wgrad(dq_buf, x2, gw_in[:E])
            _wgrad(dkv_buf, e2, gw_in[E:])
            ops.colsum_rows(dq_buf, gb_in[:E])
            ops.colsum_rows(dkv_buf, gb_in[E:])
            dxn = ops.gemm_bf16(dq_buf, pk.bwd("in_q", w_in[:E]), out_dtype=torch.float32)      # [E, E]
            den = ops.gemm_bf16(dkv_buf, pk.bwd("in_kv", w_in[E:]), out_dtype=torch.float32)   # [E, 2E]
            genc = torch.empty_like(enc)
            glne_w, glne_b = torch.empty_like(lne_w), torch.empty_like(lne_w)
            ops.layernorm_bwd(den, enc.contiguous(), lne_w, emean, erstd, genc, dgamma=glne_w, dbeta=glne_b)
        else:
            _wgrad(dq_buf, x2, gw_in)
            ops.colsum_rows(dq_buf, gb_in)
            dxn = ops.gemm_bf16(dq_buf, pk.bwd("in", w_in), out_dtype=torch.float32)           # [E, 3E]
        ops.copy2d_f32(g.view(b * S, E), dxn, accumulate=True)             # residual branch
        gx = torch.empty_like(x)
        gln_w, gln_b = torch.empty_like(ln_w), torch.empty_like(ln_w)
        ops.layernorm_bwd(dxn, x.contiguous(), ln_w, mean, rstd, gx, dgamma=gln_w, dbeta=gln_b)
        return None, gx, genc, gln_w, gln_b, glne_w, glne_b, gw_in, gb_in, gw_out, gb_out


class _FeedForwardFn(torch.autograd.Function):
    """fc2(gelu(fc1(LN(x))))   (rovr/common_layers.py:89-92)."""

    @staticmethod
    def forward(ctx, blk, x, ln_w, ln_b, w1, b1, w2, b2):
        shape = x.shape
        E = shape[-1]
        pk = blk._packed
        _, xn16, mean, rstd = ops.layernorm_fwd(x.contiguous(), ln_w, ln_b, blk.layer_norm.eps, want_f32=False)
        x2 = xn16.view(-1, E)
        h = ops.gemm_bf16(x2, pk.fwd("fc1", w1), b1)
        a = ops.gelu_fwd(h)
        drop = _draw_dropout(blk, blk.dropout.p, 1)
        if drop is not None:
            a = ops.dropout(a, *drop)                                      # nn.Dropout after the GELU (:91)
        y = ops.gemm_bf16(a, pk.fwd("fc2", w2), b2, out_dtype=torch.float32)
        ctx.blk, ctx.drop = blk, drop
        ctx.save_for_backward(x, ln_w, w1, w2, mean, rstd, x2, h, a)
        return y.view(shape)

    @staticmethod
    def backward(ctx, g):
        x, ln_w, w1, w2, mean, rstd, x2, h, a = ctx.saved_tensors
        pk = ctx.blk._packed
        E = x.shape[-1]
        g16 = ops.cast_bf16(g.contiguous().float().view(-1, E))
        gw2, gb2 = _linear_grads(g16, a, w2)
        da = ops.gemm_bf16(g16, pk.bwd("fc2", w2))
        if ctx.drop is not None:
            ops.dropout(da, *ctx.drop, out=da)
        dh = ops.gelu_bwd(da, h)
        gw1, gb1 = _linear_grads(dh, x2, w1)
        dxn = ops.gemm_bf16(dh, pk.bwd("fc1", w1), out_dtype=torch.float32)
        gx = torch.empty_like(x)
        gln_w, gln_b = torch.empty_like(ln_w), torch.empty_like(ln_w)
        ops.layernorm_bwd(dxn, x.contiguous(), ln_w, mean, rstd, gx, dgamma=gln_w, dbeta=gln_b)
        return None, gx, gln_w, gln_b, gw1, gb1, gw2, gb2


class _AddFn(torch.autograd.Function):
    """fp32 residual add (x = x + block(x), rovr/common_layers.py:101-102,113-115)."""

    @staticmethod
    def forward(ctx, a, b):
        return ops.add_f32(a.contiguous(), b.contiguous())

    @staticmethod
    def backward(ctx, g):
        return g, g


class _PosEncFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, n1, w1, b1, w2, b2):
        ctx.n1, ctx.two = n1, w2 is not None
        return ops.posenc_add(x.contiguous().float(), w1.reshape(-1), b1, n1,
                              None if w2 is None else w2.reshape(-1), b2)

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous().float()
        gw1, gb1 = ops.posenc_grad(g, ctx.n1, 0)
        gw2 = gb2 = None
        if ctx.two:
            gw2, gb2 = ops.posenc_grad(g, ctx.n1, 1)
            gw2 = gw2.view(-1, 1)
        return g, None, gw1.view(-1, 1), gb1, gw2, gb2


def _draw_dropout(blk, p, site):
    """None when dropout is inactive (eval mode or p == 0); else (p, snapshot of the block's {seed, counter}
    state, site) — and the live counter moves on, so the next forward draws another mask."""
    if not blk.training or p <= 0:
        return None
    st = getattr(blk, "_dropout_state", None)
    if st is None or st.device != next(blk.parameters()).device:
        seed = int(torch.empty((), dtype=torch.int64).random_())          # torch's CPU generator: torch.manual_seed applies
        st = torch.tensor([seed, 0], dtype=torch.int64, device=next(blk.parameters()).device)
        blk._dropout_state = st
    snap = st.clone()
    ops.dropout_advance(st)
    return (float(p), snap, site)


class ImagePositionalEncoding(nn.Module):
    """x + Linear(1, P^2 C)(arange(n^2))  (rovr/common_layers.py:7-25)."""

    def __init__(self, num_image_patches, patch_size, num_channels):
        super(ImagePositionalEncoding, self).__init__()
        self.num_image_patches = num_image_patches
        self.patch_size = patch_size
        self.num_channels = num_channels
        self.positional_encoder = nn.Linear(1, self.patch_size ** 2 * self.num_channels)

    def forward(self, x):
        _need_cuda(x, "ImagePositionalEncoding")
        n = self.num_image_patches ** 2
        if x.shape[1] != n:
            raise RuntimeError(f"expected {n} image tokens, got {x.shape[1]}")
        return _PosEncFn.apply(x, n, self.positional_encoder.weight, self.positional_encoder.bias, None, None)


class ContextPositionalEncoding(nn.Module):
    """x + patch_pos(arange(p^2)) + context_pos(arange(n))  (rovr/common_layers.py:27-52)."""

    def __init__(self, num_context_patches, patch_size, num_channels, num_context):
        super(ContextPositionalEncoding, self).__init__()
        self.num_context_patches = num_context_patches
        self.patch_size = patch_size
        self.num_channels = num_channels
        self.num_context = num_context
        self.patch_positional_encoder = nn.Linear(1, self.patch_size ** 2 * self.num_channels)
        self.context_positional_encoder = nn.Linear(1, self.patch_size ** 2 * self.num_channels)

    def forward(self, x):
        _need_cuda(x, "ContextPositionalEncoding")
        p2 = self.num_context_patches ** 2
        if x.shape[1] != p2 * self.num_context:
            raise RuntimeError(f"expected {p2 * self.num_context} context tokens, got {x.shape[1]}")
        return _PosEncFn.apply(x, p2, self.patch_positional_encoder.weight, self.patch_positional_encoder.bias,
                               self.context_positional_encoder.weight, self.context_positional_encoder.bias)


class SelfAttentionBlock(nn.Module):
    def __init__(self, hidden_dim, num_heads, dropout):
        super(SelfAttentionBlock, self).__init__()
        _check_dims(hidden_dim, num_heads)
        self.attention = nn.MultiheadAttention(hidden_dim, num_heads, dropout=dropout, batch_first=True)
        self.layer_norm = nn.LayerNorm(hidden_dim)
        self._packed = _Packed()

    def forward(self, x):
        _need_cuda(x, "SelfAttentionBlock")
        a = self.attention
        return _AttentionFn.apply(self, x.float(), None, self.layer_norm.weight, self.layer_norm.bias, None, None,
                                  a.in_proj_weight, a.in_proj_bias, a.out_proj.weight, a.out_proj.bias)


class CrossAttentionBlock(nn.Module):
    def __init__(self, hidden_dim, num_heads, dropout):
        super(CrossAttentionBlock, self).__init__()
        _check_dims(hidden_dim, num_heads)
        self.attention = nn.MultiheadAttention(hidden_dim, num_heads, dropout=dropout, batch_first=True)
        self.layer_norm = nn.LayerNorm(hidden_dim)
        self.layer_norm_encoder_output = nn.LayerNorm(hidden_dim)
        self._packed = _Packed()

    def forward(self, x, encoder_output):
        _need_cuda(x, "CrossAttentionBlock")
        a = self.attention
        return _AttentionFn.apply(self, x.float(), encoder_output.float(), self.layer_norm.weight,
                                  self.layer_norm.bias, self.layer_norm_encoder_output.weight,
                                  self.layer_norm_encoder_output.bias, a.in_proj_weight, a.in_proj_bias,
                                  a.out_proj.weight, a.out_proj.bias)


class FeedForwardBlock(nn.Module):
    def __init__(self, hidden_dim, dropout):
        super(FeedForwardBlock, self).__init__()
        if hidden_dim % 64:
            raise ValueError("B200 feed-forward path needs hidden_dim to be a multiple of 64")
        self.fc1 = nn.Linear(hidden_dim, hidden_dim // 4)
        self.fc2 = nn.Linear(hidden_dim // 4, hidden_dim)
        self.layer_norm = nn.LayerNorm(hidden_dim)
        self.dropout = nn.Dropout(dropout)
        self._packed = _Packed()

    def forward(self, x):
        _need_cuda(x, "FeedForwardBlock")
        return _FeedForwardFn.apply(self, x.float(), self.layer_norm.weight, self.layer_norm.bias, self.fc1.weight,
                                    self.fc1.bias, self.fc2.weight, self.fc2.bias)


class EncoderBlock(nn.Module):
    def __init__(self, hidden_dim, num_heads, dropout):
        super(EncoderBlock, self).__init__()
        self.attention = SelfAttentionBlock(hidden_dim, num_heads, dropout)
        self.feed_forward = FeedForwardBlock(hidden_dim, dropout)

    def forward(self, x):
        x = _AddFn.apply(x.float(), self.attention(x))
        x = _AddFn.apply(x, self.feed_forward(x))
        return x


class DecoderBlock(nn.Module):
    def __init__(self, hidden_dim, num_heads, dropout):
        super(DecoderBlock, self).__init__()
        self.attention = SelfAttentionBlock(hidden_dim, num_heads, dropout)
        self.cross_attention = CrossAttentionBlock(hidden_dim, num_heads, dropout)
        self.feed_forward = FeedForwardBlock(hidden_dim, dropout)

    def forward(self, x, encoder_output):
        x = _AddFn.apply(x.float(), self.attention(x))
        x = _AddFn.apply(x, self.cross_attention(x, encoder_output))
        x = _AddFn.apply(x, self.feed_forward(x))
        return x
