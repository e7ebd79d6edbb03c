"""Attention / feed-forward dropout in training mode (rovr/common_layers.py:58,70,87,91): the mask is a
counter-based hash recomputed in backward. Checked: keep rate and scaling of the kernel; determinism of
(state, site); EncoderBlock forward + every gradient against the oracle fed with the SAME masks; a new mask
per forward; eval mode = no dropout; graph replays draw fresh masks."""
import math

import pytest
import torch

import rovr_oracle as O

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def _dev():
    import _native
    _native.require_device()
    return torch.device("cuda:0")


def _rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-20)).item()


def test_dropout_kernel_statistics_and_determinism():
    import ops
    dev = _dev()
    st = torch.tensor([1234567, 3], dtype=torch.int64, device=dev)
    x = torch.ones(1 << 20, device=dev)
    for p in (0.1, 0.5):
        y = ops.dropout(x, p, st, 0)
        keep = (y != 0).float().mean().item()
        assert abs(keep - (1 - p)) < 3e-3, (p, keep)
        assert torch.allclose(y[y != 0], torch.full((1,), 1 / (1 - p), device=dev))
        assert torch.equal(y, ops.dropout(x, p, st, 0))                       # same (state, site): same mask
        assert not torch.equal(y, ops.dropout(x, p, st, 1))                   # another site
        st2 = st.clone()
        ops.dropout_advance(st2)
        assert int(st2[1]) == 4 and not torch.equal(y, ops.dropout(x, p, st2, 0))
        yb = ops.dropout(x.to(BF), p, st, 0)
        assert torch.equal((yb != 0), (y != 0))                              # dtype does not change the mask
    # neighbouring elements are uncorrelated
    y = ops.dropout(x, 0.5, st, 0)
    k = (y != 0).float()
    assert abs(((k[1:] * k[:-1]).mean() - 0.25).item()) < 5e-3


def test_encoder_block_training_dropout_matches_oracle_with_same_masks():
    import ops
    from common_layers import EncoderBlock
    dev = _dev()
    E, heads, B, S, p = 128, 4, 3, 32, 0.25
    torch.manual_seed(5)
    blk = EncoderBlock(E, heads, p).to(dev).train()
    sd = {k: v.detach().cpu().clone() for k, v in blk.state_dict().items()}
    x = torch.randn((B, S, E), generator=torch.Generator().manual_seed(6))
    xg = x.to(dev).requires_grad_(True)
    y = blk(xg)
    (y ** 2).mean().backward()
    # the masks this forward used: state snapshots are {seed, counter = 0} of each sub-block
    st_a = blk.attention._dropout_state.clone(); st_a[1] -= 1
    st_f = blk.feed_forward._dropout_state.clone(); st_f[1] -= 1
    t_pad = ops.pad16(S)
    m_att = ops.dropout(torch.ones((B, heads, S, t_pad), device=dev), p, st_a, 0)[..., :S].cpu()
    m_ffn = ops.dropout(torch.ones((B * S, E // 4), device=dev), p, st_f, 1).view(B, S, E // 4).cpu()
    leaf = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xr = x.clone().requires_grad_(True)
    h = xr + O.self_attention_block(leaf, "attention.", xr, heads, drop_mask=m_att)
    yr = h + O.feed_forward_block(leaf, "feed_forward.", h, drop_mask=m_ffn)
    (yr ** 2).mean().backward()
    assert _rel(y, yr) < 2e-2, _rel(y, yr)
    assert _rel(xg.grad, xr.grad) < 2e-2
    for n, prm in blk.named_parameters():
        assert _rel(prm.grad, leaf[n].grad) < 2e-2, (n, _rel(prm.grad, leaf[n].grad))
    # dropout really happened, a second forward draws another mask, eval mode is deterministic and mask-free
    y_nodrop = xr.detach() + O.self_attention_block(sd, "attention.", xr.detach(), heads)
    y_nodrop = y_nodrop + O.feed_forward_block(sd, "feed_forward.", y_nodrop)
    assert _rel(y, y_nodrop) > 5e-2
    y2 = blk(xg)
    assert not torch.equal(y2, y)
    blk.eval()
    ye = blk(xg)
    assert torch.equal(ye, blk(xg)) and _rel(ye, y_nodrop) < 2e-2


def test_graph_replays_draw_fresh_dropout_masks():
    from common_layers import FeedForwardBlock
    from graphs import GraphedFunction
    dev = _dev()
    torch.manual_seed(7)
    blk = FeedForwardBlock(256, 0.5).to(dev).train()
    x = torch.randn((4, 16, 256), device=dev)

    def fn(xx):
        out = blk(xx)
        (out ** 2).mean().backward()
        return out
    g = GraphedFunction(fn, (x,), modules=[blk])
    a = g(x).clone()
    b = g(x).clone()
    assert not torch.equal(a, b), "a replay must advance the dropout counter"
