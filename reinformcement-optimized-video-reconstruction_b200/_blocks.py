"""Building blocks shared by the policy-network drop-ins: conv / transposed-conv / 1x1-conv +
train-mode BatchNorm + ReLU, forward and hand-written backward over the C ABI (ops.py).

A network's trunk is executed eagerly inside ONE torch.autograd.Function (like LocalNet): the
forward helpers below launch the kernels and return a small record of what backward needs; the
backward helpers consume the record. No torch operator computes anything here — torch only
allocates buffers.
"""
import torch

import ops

BF = torch.bfloat16


class PackedWeights:
    """bf16 GEMM-operand copies of fp32 master weights, refreshed when a parameter changes."""

    def __init__(self):
        self._cache = {}

    def get(self, key, param, builder):
        ver = (param.data_ptr(), param._version)
        hit = self._cache.get(key)
        if hit is not None and hit[0] == ver:
            return hit[1]
        val = builder(param.detach())
        self._cache[key] = (ver, val)
        return val


class GradArena:
    """One flat fp32 buffer holding every parameter gradient a trunk's backward produces (views are handed
    out in backward-completion order). Data-parallel runs all-reduce the whole buffer with ONE collective
    (data_parallel.GradientAverager finds it through the gradients' shared storage)."""

    def __init__(self, P):
        first = next(iter(P.values()))
        self.flat = torch.empty(sum(p.numel() for p in P.values()), dtype=torch.float32, device=first.device)
        self.off = 0

    def take(self, p, zero=False):
        n = p.numel()
        v = self.flat[self.off:self.off + n].view(p.shape)
        self.off += n
        return v.zero_() if zero else v


def padded_vector(v, n):
    """fp32 vector zero-padded to n entries (bias of a layer whose Cout is padded to 16)."""
    if v.numel() == n:
        return v
    out = torch.zeros(n, dtype=torch.float32, device=v.device)
    ops.copy2d_f32(v.detach().reshape(1, -1), out[: v.numel()].reshape(1, -1))
    return out


class TrunkOps:
    """Forward / backward helpers bound to one network's parameters.

    P: {name: parameter tensor}; B: {name: buffer tensor} (BatchNorm running stats);
    G: {name: gradient tensor} filled by the backward helpers; packed: PackedWeights.
    """

    def __init__(self, P, B, packed, training=True, eps=1e-5, momentum=0.1):
        self.P, self.B, self.packed = P, B, packed
        self.training = training
        self.eps, self.momentum = eps, momentum
        self.G = {}
        self.arena = None

    # -- helpers ---------------------------------------------------------------------------------
    def _grad(self, name, zero=False):
        if self.arena is None:
            self.arena = GradArena(self.P)
        g = self.arena.take(self.P[name], zero)
        self.G[name] = g
        return g

    def _bn_fwd(self, bn, raw, y, c_valid):
        gamma, beta = self.P[bn + ".weight"], self.P[bn + ".bias"]
        if self.training:
            rm, rv, nbt = self.B[bn + ".running_mean"], self.B[bn + ".running_var"], self.B[bn + ".num_batches_tracked"]
            return ops.bn_train_fwd(raw, y, gamma, beta, c_valid, self.eps, self.momentum, rm, rv, nbt, relu=True)
        # module.eval(): running statistics, buffers untouched (the reference's drivers never do this
        # with the policy networks, but a validation loop would)
        rm, rv = self.B[bn + ".running_mean"], self.B[bn + ".running_var"]
        rstd = ops.bn_eval_fwd(raw, y, gamma, beta, c_valid, self.eps, rm, rv, relu=True)
        return rm, rstd

    def _bias_grad(self, name, draw, c_valid):
        """A bias that feeds train-mode BatchNorm has an exactly-zero gradient (the batch mean is
        subtracted right after it); autograd in the reference produces fp32 rounding noise around
        zero there, and column sums of the bf16 gradient would only add rounding noise of their own.
        In eval mode the statistics are constants and the gradient is the column sum of draw."""
        p = self.P[name]
        if self.training:
            self._grad(name, zero=True)
            return
        full = torch.empty(draw.shape[-1], dtype=torch.float32, device=p.device)
        ops.colsum(draw, full)
        g = self._grad(name)
        ops.copy2d_f32(full[:c_valid].reshape(1, -1), g.reshape(1, -1))

    def _bn_bwd(self, bn, gy, y, raw, mean, rstd, c_valid):
        draw = torch.empty_like(raw)
        ops.bn_train_bwd(gy, y, raw, draw, self.P[bn + ".weight"], mean, rstd, c_valid,
                         self._grad(bn + ".weight"), self._grad(bn + ".bias"), relu=True, eval_mode=not self.training)
        return draw

    # -- Conv2d 3x3 + BN + ReLU --------------------------------------------------------------------
    def cbr3_fwd(self, conv, bn, x, y):
        """y (NHWC bf16 view, possibly a concat slice) = relu(bn(conv3x3(x))). Returns the record."""
        w = self.P[conv + ".weight"]
        cout = w.shape[0]
        wk = self.packed.get((conv, "f"), w, lambda t: ops.repack_conv3x3(t, False))
        raw = torch.empty(y.shape, dtype=BF, device=x.device)
        ops.conv3x3_fprop(x, wk, self.P[conv + ".bias"], raw, relu=False)
        mean, rstd = self._bn_fwd(bn, raw, y, cout)
        return (raw, mean, rstd)

    def cbr3_bwd(self, conv, bn, rec, x, y, gy, gx):
        """gy = dL/dy. Fills the parameter gradients; if gx is given, writes dL/dx into it."""
        raw, mean, rstd = rec
        w = self.P[conv + ".weight"]
        draw = self._bn_bwd(bn, gy, y, raw, mean, rstd, w.shape[0])
        ops.conv3x3_wgrad(draw, x, self._grad(conv + ".weight"))
        self._bias_grad(conv + ".bias", draw, w.shape[0])
        if gx is not None:
            wd = self.packed.get((conv, "d"), w, lambda t: ops.repack_conv3x3(t, True))
            ops.conv3x3_dgrad(draw, wd, gx)

    # -- ConvTranspose2d 2x2 s2 + BN + ReLU ----------------------------------------------------------
    def ubr_fwd(self, up, bn, x, y):
        w = self.P[up + ".weight"]
        cout = w.shape[1]
        wk = self.packed.get((up, "f"), w, lambda t: ops.repack_convT2x2(t, False))
        raw = torch.empty(y.shape, dtype=BF, device=x.device)
        ops.convT2x2_fprop(x, wk, self.P[up + ".bias"], raw, relu=False)
        mean, rstd = self._bn_fwd(bn, raw, y, cout)
        return (raw, mean, rstd)

    def ubr_bwd(self, up, bn, rec, x, y, gy, gx):
        raw, mean, rstd = rec
        w = self.P[up + ".weight"]
        draw = self._bn_bwd(bn, gy, y, raw, mean, rstd, w.shape[1])
        ops.convT2x2_wgrad(draw, x, self._grad(up + ".weight"))
        self._bias_grad(up + ".bias", draw, w.shape[1])
        if gx is not None:
            wd = self.packed.get((up, "d"), w, lambda t: ops.repack_convT2x2(t, True))
            ops.convT2x2_dgrad(draw, wd, gx)

    # -- Conv2d 1x1 + BN + ReLU (channel counts padded to 16) -----------------------------------------
    def cbr1_fwd(self, conv, bn, x, y):
        w = self.P[conv + ".weight"]
        cout = w.shape[0]
        Bn, H, W, cin_pad = x.shape
        cpad = y.shape[3]
        wk = self.packed.get((conv, "f"), w, lambda t: ops.repack_linear(t, False))
        bias = self.packed.get((conv, "b"), self.P[conv + ".bias"], lambda t: padded_vector(t, cpad))
        raw = torch.empty(y.shape, dtype=BF, device=x.device)
        ops.gemm_bf16(x.reshape(-1, cin_pad), wk, bias, out=raw.reshape(-1, cpad))
        mean, rstd = self._bn_fwd(bn, raw, y, cout)
        return (raw, mean, rstd)

    def cbr1_bwd(self, conv, bn, rec, x, y, gy, gx):
        raw, mean, rstd = rec
        w = self.P[conv + ".weight"]
        cout, cin = w.shape[0], w.shape[1]
        cpad, cin_pad = y.shape[3], x.shape[3]
        draw = self._bn_bwd(bn, gy, y, raw, mean, rstd, cout)
        d2, x2 = draw.reshape(-1, cpad), x.reshape(-1, cin_pad)
        ops.gemm_wgrad(d2, x2, self._grad(conv + ".weight"))
        self._bias_grad(conv + ".bias", draw, cout)
        if gx is not None:
            wd = self.packed.get((conv, "d"), w, lambda t: ops.repack_linear(t, True))
            ops.gemm_bf16(d2, wd, None, out=gx.reshape(-1, cin_pad))


# ------------------------------------------------------------------------------------------------
# Emulated-fp32 trunks (csrc/fp32x.cuh): fp32 NHWC activations and gradients, convolutions as
# split-bf16 products stacked along the contraction dimension. Default for the policy networks:
# their gradients pass through 8x8 / 4x4 / 2x2 max-pools whose arg-max a bf16-operand convolution
# gets wrong in a few percent of the windows (0.2-0.6 L2-rel gradient error, scripts/exp/
# precision_policy.py); with 6-term forward / 3-term backward products outputs and gradients match
# the fp32 reference to ~1e-4.
# ------------------------------------------------------------------------------------------------
F32 = torch.float32


def pad16(c):
    return (c + 15) // 16 * 16


class TrunkOpsF32:
    """Forward / backward helpers of an emulated-fp32 trunk bound to one network's parameters.

    Activations: fp32 NHWC views. The operand of a convolution is `stack6(x)`: bf16
    [B, H, W, 6 * cb] with cb = pad8(C) (blocks [h m l h m h]); gradients are stacked as
    [h m h] (3 * pad16(Cout) channels). A layer record keeps (operand stack, raw conv output,
    mean, rstd)."""

    def __init__(self, P, B, packed, training=True, eps=1e-5, momentum=0.1):
        self.P, self.B, self.packed = P, B, packed
        self.training = training
        self.eps, self.momentum = eps, momentum
        self.G = {}
        self.arena = None

    # -- helpers ---------------------------------------------------------------------------------
    def _grad(self, name, zero=False):
        if self.arena is None:
            self.arena = GradArena(self.P)
        g = self.arena.take(self.P[name], zero)
        self.G[name] = g
        return g

    def stack6(self, x, pool=None):
        """fp32 NHWC view -> forward operand; pool = (kh, kw, sh, sw) fuses nn.MaxPool2d in front."""
        return ops.split_stack(x, 6, ops.pad8(x.shape[3]), pool=pool)

    def _bn_fwd(self, bn, raw, y, c_valid):
        gamma, beta = self.P[bn + ".weight"], self.P[bn + ".bias"]
        if self.training:
            rm, rv, nbt = self.B[bn + ".running_mean"], self.B[bn + ".running_var"], self.B[bn + ".num_batches_tracked"]
            return ops.bn_f32_train_fwd(raw, y, gamma, beta, c_valid, self.eps, self.momentum, rm, rv, nbt, relu=True)
        rm, rv = self.B[bn + ".running_mean"], self.B[bn + ".running_var"]
        return rm, ops.bn_f32_eval_fwd(raw, y, gamma, beta, c_valid, self.eps, rm, rv, relu=True)

    def _bn_bwd(self, bn, gy, y, raw, mean, rstd, c_valid):
        draw = torch.empty(raw.shape, dtype=F32, device=raw.device)
        ops.bn_f32_bwd(gy, y, raw, draw, self.P[bn + ".weight"], mean, rstd, c_valid, self._grad(bn + ".weight"),
                       self._grad(bn + ".bias"), relu=True, eval_mode=not self.training)
        return draw

    def _bias_grad(self, name, draw, c_valid):
        """Exactly zero in train mode (the batch mean is subtracted right after the bias, see
        TrunkOps._bias_grad); the column sums of draw in eval mode."""
        p = self.P[name]
        if self.training:
            self._grad(name, zero=True)
            return
        full = torch.empty(draw.shape[-1], dtype=F32, device=p.device)
        ops.colsum_f32(draw, full)
        ops.copy2d_f32(full[:c_valid].reshape(1, -1), self._grad(name).reshape(1, -1))

    def _wk(self, name, tag, w, stack_dim, nterms, cb, repack):
        """bf16 operand of the stacked product: split the fp32 weight into pieces, then the usual re-pack."""
        return self.packed.get((name, tag), w, lambda t: repack(ops.split_weights(t, stack_dim, nterms, cb)))

    # -- Conv2d 3x3 + BN + ReLU --------------------------------------------------------------------
    def cbr3_fwd(self, conv, bn, xs, y):
        """xs: stacked operand of the input; y: fp32 NHWC view (possibly a concat slice) that receives
        relu(bn(conv(x)))."""
        w = self.P[conv + ".weight"]
        cout, cin = w.shape[0], w.shape[1]
        wk = self._wk(conv, "f6", w, 1, 6, ops.pad8(cin), lambda t: ops.repack_conv3x3(t, False))
        raw = torch.empty(y.shape, dtype=F32, device=xs.device)
        ops.conv3x3_f32out(xs, wk, self.P[conv + ".bias"], raw)
        mean, rstd = self._bn_fwd(bn, raw, y, cout)
        return (xs, raw, mean, rstd)

    def cbr3_bwd(self, conv, bn, rec, y, gy, gx):
        xs, raw, mean, rstd = rec
        w = self.P[conv + ".weight"]
        cout, cin = w.shape[0], w.shape[1]
        cbi, cbo = ops.pad8(cin), pad16(cout)
        draw = self._bn_bwd(bn, gy, y, raw, mean, rstd, cout)
        ds = ops.split_stack(draw, 3, cbo)
        dwp = torch.empty((2 * cbo, 2 * cbi, 3, 3), dtype=F32, device=w.device)
        ops.conv3x3_wgrad(ds[..., :2 * cbo], xs[..., :2 * cbi], dwp)
        ops.blocksum4(dwp, self._grad(conv + ".weight"), cbo, cbi)
        self._bias_grad(conv + ".bias", draw, cout)
        if gx is not None:
            wd = self._wk(conv, "d3", w, 0, 3, cbo, lambda t: ops.repack_conv3x3(t, True))
            ops.conv3x3_f32out(ds, wd, None, gx)

    # -- ConvTranspose2d 2x2 s2 + BN + ReLU ----------------------------------------------------------
    def ubr_fwd(self, up, bn, xs, y):
        w = self.P[up + ".weight"]
        cin, cout = w.shape[0], w.shape[1]
        wk = self._wk(up, "f6", w, 0, 6, ops.pad8(cin), lambda t: ops.repack_convT2x2(t, False))
        raw = torch.empty(y.shape, dtype=F32, device=xs.device)
        ops.convT2x2_fprop_f32out(xs, wk, self.P[up + ".bias"], raw)
        mean, rstd = self._bn_fwd(bn, raw, y, cout)
        return (xs, raw, mean, rstd)

    def ubr_bwd(self, up, bn, rec, y, gy, gx):
        xs, raw, mean, rstd = rec
        w = self.P[up + ".weight"]
        cin, cout = w.shape[0], w.shape[1]
        cbi, cbo = ops.pad8(cin), pad16(cout)
        draw = self._bn_bwd(bn, gy, y, raw, mean, rstd, cout)
        ds = ops.split_stack(draw, 3, cbo)
        dwp = torch.empty((2 * cbi, 2 * cbo, 2, 2), dtype=F32, device=w.device)
        ops.convT2x2_wgrad(ds[..., :2 * cbo], xs[..., :2 * cbi], dwp)
        ops.blocksum4(dwp, self._grad(up + ".weight"), cbi, cbo)
        self._bias_grad(up + ".bias", draw, cout)
        if gx is not None:
            wd = self._wk(up, "d3", w, 1, 3, cbo, lambda t: ops.repack_convT2x2(t, True))
            ops.convT2x2_dgrad_f32out(ds, wd, gx)

    # -- Conv2d 1x1 + BN + ReLU (output channels padded to 16) ---------------------------------------
    def cbr1_fwd(self, conv, bn, xs, y):
        w = self.P[conv + ".weight"]
        cout, cin = w.shape[0], w.shape[1]
        cpad = y.shape[3]
        wk = self._wk(conv, "f6", w.reshape(cout, cin), 1, 6, ops.pad8(cin), lambda t: ops.repack_linear(t, False))
        bias = self.packed.get((conv, "b"), self.P[conv + ".bias"], lambda t: padded_vector(t, cpad))
        raw = torch.empty(y.shape, dtype=F32, device=xs.device)
        ops.gemm_bf16(xs.reshape(-1, xs.shape[3]), wk, bias, out=raw.reshape(-1, cpad))
        mean, rstd = self._bn_fwd(bn, raw, y, cout)
        return (xs, raw, mean, rstd)

    def cbr1_bwd(self, conv, bn, rec, y, gy, gx):
        xs, raw, mean, rstd = rec
        w = self.P[conv + ".weight"]
        cout, cin = w.shape[0], w.shape[1]
        cpad = y.shape[3]
        cbi = ops.pad8(cin)
        draw = self._bn_bwd(bn, gy, y, raw, mean, rstd, cout)
        ds = ops.split_stack(draw, 3, cpad)                                   # [.., 3 * cpad]
        d2, x2 = ds.reshape(-1, 3 * cpad), xs.reshape(-1, xs.shape[3])
        dwp = torch.empty((2 * cpad, 2 * cbi), dtype=F32, device=w.device)
        ops.gemm_wgrad(d2[:, :2 * cpad], x2[:, :2 * cbi], dwp)
        ops.blocksum4(dwp, self._grad(conv + ".weight").view(cout, cin), cpad, cbi)
        self._bias_grad(conv + ".bias", draw, cout)
        if gx is not None:
            wd = self._wk(conv, "d3", w.reshape(cout, cin), 0, 3, cpad, lambda t: ops.repack_linear(t, True))
            ops.gemm_bf16(d2, wd, None, out=gx.reshape(-1, gx.shape[3]))
