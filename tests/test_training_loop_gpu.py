"""the reference's training loop (rovr/train_local_net_unet.py:71,102-116 —
`optimizer.zero_grad(); y = net(x, ctx); loss = mse(y, t); loss.backward(); optimizer.step()` with
torch.optim.Adam, lr 1e-4) stepped over the drop-in, eager and through GraphedTrainingStep, against
the same loop over the CPU fp32 oracle. Adam stays torch.optim (not a kernel target); what is
checked here is that the drop-in's gradients arrive where the optimizer looks for them —
including after the loop's own `zero_grad()` (set_to_none) — and that the bf16 operand copies
follow the updated fp32 masters."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

STEPS, LR = 20, 1e-4        # rovr/train_local_net_unet.py:71 (lr), 20 steps of its loop
TOL = 2e-2


def _oracle_loop(sd, batches):
    import rovr_oracle as O
    leaf = {k: (v.clone().requires_grad_(True) if k in O.LOCALNET_LIVE else v.clone()) for k, v in sd.items()}
    opt = torch.optim.Adam([leaf[k] for k in O.LOCALNET_LIVE], lr=LR)
    losses = []
    for x, c, t in batches:
        opt.zero_grad()
        loss = F.mse_loss(O.localnet_forward(leaf, x, c), t)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    return losses, {k: leaf[k].detach() for k in O.LOCALNET_LIVE}


def _rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-20)).item()


@pytest.mark.parametrize("mode", ["eager", "graphed"])
def test_adam_loop_matches_oracle(mode):
    import _native
    import rovr_oracle as O
    from local_net import GraphedTrainingStep, LocalNetworkUNetNorm
    _native.require_device()
    dev = torch.device("cuda:0")
    sd = O.localnet_state_dict(0)
    batches = [O.synthetic_localnet_batch(4, 64, 64, seed=300 + (i % 4)) for i in range(STEPS)]
    ref_losses, ref_w = _oracle_loop(sd, batches)

    net = LocalNetworkUNetNorm()
    net.load_state_dict(sd, strict=True)
    net = net.to(dev)
    opt = torch.optim.Adam(net.parameters(), lr=LR)          # all 72-key parameters, like the reference
    w0 = {n: p.detach().clone() for n, p in net.named_parameters()}
    step = None
    if mode == "graphed":
        step = GraphedTrainingStep(net, *[v.to(dev) for v in batches[0]])
        assert step.grads_alias_buckets()
    losses = []
    for x, c, t in batches:
        x, c, t = x.to(dev), c.to(dev), t.to(dev)
        opt.zero_grad()                                      # set_to_none=True in torch 2.x
        if step is None:
            loss = F.mse_loss(net(x, c), t)                  # the reference's own lines :105-107
            loss.backward()
        else:
            loss = step(x, c, t)
            assert step.grads_alias_buckets()                # .grad re-attached to the static buckets
        opt.step()
        losses.append(float(loss))
    named = dict(net.named_parameters())
    for i, (a, b) in enumerate(zip(losses, ref_losses)):
        assert abs(a - b) <= TOL * abs(b), f"step {i}: loss {a} vs oracle {b}"
    assert losses[-1] < losses[0]
    for k, wr in ref_w.items():
        assert not torch.equal(named[k].detach(), w0[k]), f"{k} was never updated"
        assert _rel(named[k], wr) < TOL, f"{k}: weights differ {_rel(named[k], wr):.3e}"
        # the update itself (w - w0): Adam divides by sqrt(v), so entries with a near-zero gradient amplify the
        # 2e-2-class gradient noise of the bf16 path into sign-level differences; a loose sanity bound only
        du, dr = named[k].detach().cpu() - w0[k].cpu(), wr - sd[k]
        assert (du - dr).norm() <= 0.35 * dr.norm(), f"{k}: Adam update differs {_rel(du, dr):.3e}"
    # the never-applied BatchNorm parameters stay without gradient, exactly as in the reference
    assert all(p.grad is None for n, p in net.named_parameters() if n.startswith("bn"))


def test_graphed_step_survives_larger_eager_call():
    """a later eager call that needs a bigger scratch workspace must not free the buffer a
    captured graph has baked in."""
    import rovr_oracle as O
    from local_net import GraphedTrainingStep, LocalNetworkUNetNorm
    dev = torch.device("cuda:0")
    net = LocalNetworkUNetNorm()
    net.load_state_dict(O.localnet_state_dict(0), strict=True)
    net = net.to(dev)
    small = [v.to(dev) for v in O.synthetic_localnet_batch(1, 32, 32, seed=7)]
    step = GraphedTrainingStep(net, *small)
    l0 = step().clone()
    g0 = {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}
    big = [v.to(dev) for v in O.synthetic_localnet_batch(6, 128, 128, seed=8)]
    _, loss = net.forward_with_mse(*big)                      # eager, much larger workspace
    loss.backward()
    junk = [torch.full((1 << 22,), 7.0, device=dev) for _ in range(8)]   # reuse whatever was freed
    net.zero_grad()
    l1 = step().clone()
    torch.cuda.synchronize()
    assert torch.equal(l0, l1)
    for n, p in net.named_parameters():
        if n in g0:
            assert torch.equal(p.grad, g0[n]), n
    del junk


def test_uint8_frames_feed_equals_host_totensor():
    """Frames shipped as uint8 (as decoded, rovr/video_ds.py:107-114) and converted on the GPU give exactly
    the step that host-side ToTensor (uint8 / 255 -> fp32, shipped as fp32) gives — through the feeder's
    device-side conversion and through GraphedTrainingStep's direct uint8 inputs."""
    import rovr_oracle as O
    from feeder import DeviceFeeder
    from local_net import GraphedTrainingStep, LocalNetworkUNetNorm
    dev = torch.device("cuda:0")
    net = LocalNetworkUNetNorm()
    net.load_state_dict(O.localnet_state_dict(0), strict=True)
    net = net.to(dev)
    u8 = [(v * 255.0).round().to(torch.uint8).pin_memory() for v in O.synthetic_localnet_batch(2, 64, 64, seed=51)]
    f32 = [v.float() / 255.0 for v in u8]                                    # torchvision ToTensor on the host
    _, want = net.forward_with_mse(*[v.to(dev) for v in f32])
    (xd, cd, td), = list(DeviceFeeder([tuple(u8)], dev))                     # converted behind the H2D copy
    assert xd.dtype == torch.float32 and torch.equal(xd.cpu(), f32[0])
    _, got = net.forward_with_mse(xd, cd, td)
    assert torch.equal(got, want)
    step = GraphedTrainingStep(net, *[v.to(dev) for v in f32])
    (xu, cu, tu), = list(DeviceFeeder([tuple(u8)], dev, to_float=False))
    assert xu.dtype == torch.uint8
    assert torch.equal(step(xu, cu, tu), want.detach())


def test_double_buffered_graph_inputs_and_side_stream_readback():
    """GraphedTrainingStep(input_sets=2) + DeviceFeeder(slots=...) + ScalarReadback(side_stream=True): six different
    uint8 batches give, in order, exactly the losses and the final gradients of six eager steps."""
    import rovr_oracle as O
    from feeder import DeviceFeeder, ScalarReadback
    from local_net import GraphedTrainingStep, LocalNetworkUNetNorm
    dev = torch.device("cuda:0")
    net = LocalNetworkUNetNorm()
    net.load_state_dict(O.localnet_state_dict(0), strict=True)
    net = net.to(dev)
    host = [tuple((v * 255.0).round().to(torch.uint8).pin_memory() for v in O.synthetic_localnet_batch(2, 64, 64, seed=60 + i))
            for i in range(6)]
    want, grads = [], None
    for xb, cb, tb in host:
        net.zero_grad(set_to_none=True)
        # host-side ToTensor (IEEE division; torch's CUDA div-by-scalar multiplies by the reciprocal instead)
        _, loss = net.forward_with_mse(*[(v.float() / 255).to(dev) for v in (xb, cb, tb)])
        loss.backward()
        want.append(float(loss))
        grads = {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}
    assert len(set(want)) == 6
    step = GraphedTrainingStep(net, *[(v.float() / 255).to(dev) for v in host[0]], input_sets=2)
    assert len(step.input_slots) == 2 and step.input_slots[0][0].data_ptr() != step.input_slots[1][0].data_ptr()
    for rep in range(2):                                   # twice: the slots / graphs are reused across loops
        rb = ScalarReadback(dev, lag=1, side_stream=True)
        got = []
        for k, _ in enumerate(DeviceFeeder(iter(host), dev, slots=step.input_slots)):
            v = rb.exchange(step.replay(k % 2))
            if v is not None:
                got.append(v)
        got.append(rb.drain())
        assert got == want, (rep, got, want)
        torch.cuda.synchronize()
        for n, p in net.named_parameters():
            if n in grads:
                assert torch.equal(p.grad, grads[n]), n
