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

    # -- helpers ---------------------------------------------------------------------------------
    def _grad(self, name):
        p = self.P[name]
        g = torch.empty(p.shape, dtype=torch.float32, device=p.device)
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
            self.G[name] = torch.zeros(p.shape, dtype=torch.float32, device=p.device)
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
