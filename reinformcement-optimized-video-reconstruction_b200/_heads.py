"""fp32 autograd Functions of the policy heads over the C ABI: nn.Linear for small batches,
standardisation, the reference's scatter + keepdim-less standardisation of frame logits, and the
gumbel-softmax log-probabilities. torch only allocates; every value is computed by librovr_b200.
"""
import torch

import ops


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


class LinearF32(torch.autograd.Function):
    """y = x @ w^T + b in fp32 (nn.Linear: rovr/policy_net_1.py:94; rovr/policy_net_2.py:79;
    rovr/resnet_extractor.py:46; rovr/action_lstm.py:35)."""

    @staticmethod
    def forward(ctx, x, w, b):
        x2 = _c(x.reshape(-1, x.shape[-1]).float())
        y = ops.linear_f32_fwd(x2, _c(w), b)
        ctx.save_for_backward(x2, w)
        ctx.has_bias = b is not None
        ctx.xshape = x.shape
        return y.reshape(*x.shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, gy):
        x2, w = ctx.saved_tensors
        g2 = _c(gy.reshape(-1, w.shape[0]).float())
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = ops.linear_f32_dgrad(g2, _c(w)).reshape(ctx.xshape)
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            gw = torch.empty_like(w)
            gb = torch.empty(w.shape[0], dtype=torch.float32, device=w.device) if ctx.has_bias else None
            ops.linear_f32_wgrad(g2, x2, gw, gb)
        return gx, gw, gb


class Standardize(torch.autograd.Function):
    """(x - mean) / (std_unbiased + eps_add) along `dim` of a 2-D tensor, mean / std with keepdim
    (rovr/policy_net_1.py:91-93 with dim=1, eps 0; rovr/policy_net_2.py:104-106 with dim=0, eps .001)."""

    @staticmethod
    def forward(ctx, x, dim, eps_add):
        y, sig = ops.standardize_fwd(_c(x.float()), dim, eps_add)
        ctx.save_for_backward(y, sig)
        ctx.dim, ctx.eps_add = dim, eps_add
        return y

    @staticmethod
    def backward(ctx, g):
        y, sig = ctx.saved_tensors
        return ops.standardize_bwd(_c(g.float()), y, sig, ctx.dim, ctx.eps_add), None, None


class MaskedLogits(torch.autograd.Function):
    """scatter 0 at `target` then, if standardize, (l - l.mean(dim=1)) / (l.std(dim=1, keepdim) + .1)
    with the reference's keepdim-less mean broadcast (rovr/policy_net_2.py:117-122,138;
    rovr/policy_net_1.py:100). The reference scatters in place on the Linear output; here the
    Linear output is copied first (its aliasing is unobservable)."""

    @staticmethod
    def forward(ctx, logits, target, standardize):
        l = torch.empty_like(logits, memory_format=torch.contiguous_format)
        ops.copy2d_f32(_c(logits), l)
        if target is not None:
            target = _c(target.to(torch.int64))
        out, sig = ops.head_mask_std_fwd(l, target, standardize)
        ctx.standardize = standardize
        ctx.target = target
        ctx.save_for_backward(l, out, sig)
        return out if standardize else l

    @staticmethod
    def backward(ctx, g):
        l, out, sig = ctx.saved_tensors
        return ops.head_mask_std_bwd(_c(g.float()), l, out, sig, ctx.target, ctx.standardize), None, None


class GumbelLogProb(torch.autograd.Function):
    """log-probability of `action` under F.gumbel_softmax(logits, tau, hard=False, dim=1) with the
    Exp(1) draw `expo` supplied: mode 3 = single index (rovr/policy_net_1.py:113-114), mode 4 = an
    ordered pair through the outer product p p^T, (log p_a0 p_a1)/2 + 0.69314
    (rovr/policy_net_2.py:139-141)."""

    @staticmethod
    def forward(ctx, logits, expo, tau, mode, action):
        action = _c(action.to(torch.int64))
        probs, _, val = ops.head_gumbel_fwd(_c(logits), expo, tau, mode, action)
        ctx.save_for_backward(probs, action)
        ctx.tau, ctx.mode = tau, mode
        return val

    @staticmethod
    def backward(ctx, g):
        probs, action = ctx.saved_tensors
        return ops.head_gumbel_bwd(probs, _c(g.float()), ctx.tau, ctx.mode, action), None, None, None, None


def exponential_like(logits):
    """The Exp(1) draw F.gumbel_softmax makes (torch/nn/functional.py: `torch.empty_like(logits)
    .exponential_()`), taken from the same global generator in the same order as the reference."""
    return torch.empty_like(logits, memory_format=torch.legacy_contiguous_format).exponential_()
