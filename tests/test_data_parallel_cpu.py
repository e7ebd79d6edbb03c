"""CPU (gloo, world_size 2): host-side logic of the data-parallel path — batch sharding, bucket
layout in backward-completion order, overlapped all-reduce + wait semantics, parameter broadcast."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                    "reinformcement-optimized-video-reconstruction_b200"))
    from data_parallel import GradientBuckets, broadcast_parameters, shard_range
    from local_net import LocalNetworkUNetNorm, _DECODER, _ENCODER, _flat_bucket

    torch.manual_seed(100 + rank)  # different init per rank on purpose
    net = LocalNetworkUNetNorm()
    broadcast_parameters(net)
    ref = [p.detach().clone() for p in net.parameters()]
    gathered = [torch.zeros_like(ref[0]) for _ in range(world)]
    dist.all_gather(gathered, ref[0])
    same_after_broadcast = all(torch.equal(g, gathered[0]) for g in gathered)

    gb = GradientBuckets(net)
    P = dict(net.named_parameters())
    dec, dviews = _flat_bucket(P, _DECODER, torch.device("cpu"))
    enc, eviews = _flat_bucket(P, _ENCODER, torch.device("cpu"))
    dec.fill_(float(rank + 1))
    enc.copy_(torch.arange(enc.numel(), dtype=torch.float32) * (rank + 1))
    net._bucket_ready(0, dec)   # decoder bucket is complete first
    net._bucket_ready(2, enc)
    net._buckets_wait()
    mean_scale = sum(r + 1 for r in range(world)) / world
    ok_dec = torch.allclose(dec, torch.full_like(dec, mean_scale))
    ok_enc = torch.allclose(enc, torch.arange(enc.numel(), dtype=torch.float32) * mean_scale)
    # views alias the flat buffers: per-parameter gradients see the averaged values
    ok_view = torch.allclose(dviews["conv8.bias"], torch.full((3,), mean_scale))
    # the post-replay path of GraphedTrainingStep: both buckets averaged in one call
    dec.fill_(float(rank + 1))
    enc.fill_(2.0 * (rank + 1))
    net._grad_bucket_reduce((dec, enc))
    ok_dec = ok_dec and torch.allclose(dec, torch.full_like(dec, mean_scale))
    ok_enc = ok_enc and torch.allclose(enc, torch.full_like(enc, 2.0 * mean_scale))
    from local_net import _BUCKETS, _make_buckets
    flats, views = _make_buckets(P, torch.device("cpu"))
    sizes = [f.numel() for f in flats]
    lo, hi = shard_range(50, rank, world)
    assert sizes == [2237507, 512 * 256 * 9 + 512, 1554432 - (512 * 256 * 9 + 512)] and len(views) == 22, sizes
    q.put((rank, same_after_broadcast, ok_dec, ok_enc, ok_view, dec.numel(), enc.numel(), gb.launched, lo, hi))
    dist.destroy_process_group()


def test_gradient_buckets_gloo_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, same, ok_dec, ok_enc, ok_view, ndec, nenc, launched, lo, hi in res:
        assert same and ok_dec and ok_enc and ok_view
        # bucket sizes of SURVEY.md §8e: decoder 2 237 507, encoder 1 554 432 elements
        assert (ndec, nenc) == (2237507, 1554432)
        assert launched == 2
    assert [(r[8], r[9]) for r in res] == [(0, 25), (25, 50)]


def _avg_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                    "reinformcement-optimized-video-reconstruction_b200"))
    from data_parallel import GradientAverager
    # a module whose first three gradients are views of one flat arena (what the trunk Functions of the
    # policy networks produce, _blocks.GradArena) and whose last two are separate tensors (the fp32 heads)
    m = torch.nn.Module()
    shapes = [(4, 3, 3, 3), (4,), (5, 4), (7, 5), (7,)]
    for i, sh in enumerate(shapes):
        m.register_parameter(f"p{i}", torch.nn.Parameter(torch.zeros(sh)))
    params = list(m.parameters())
    arena = torch.empty(sum(p.numel() for p in params[:3]))
    off = 0
    for i, p in enumerate(params):
        val = float((rank + 1) * (i + 1))
        if i < 3:
            p.grad = arena[off:off + p.numel()].view(p.shape)
            off += p.numel()
            p.grad.fill_(val)
        else:
            p.grad = torch.full(p.shape, val)
    frozen = torch.nn.Parameter(torch.zeros(3))
    m.register_parameter("frozen", frozen)                  # no gradient: must be skipped
    avg = GradientAverager(m)
    avg.average()
    mean = sum(r + 1 for r in range(world)) / world
    ok = all(torch.allclose(p.grad, torch.full(p.shape, mean * (i + 1))) for i, p in enumerate(params))
    q.put((rank, ok, avg.collectives, frozen.grad is None))
    dist.destroy_process_group()


def test_gradient_averager_gloo_world2():
    """Generic gradient averaging (policy networks): one collective per shared gradient arena + one for the
    loose head gradients."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_avg_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, collectives, frozen_none in res:
        assert ok and frozen_none
        assert collectives == 2


def test_shard_range_covers_everything():
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                    "reinformcement-optimized-video-reconstruction_b200"))
    from data_parallel import shard_range
    for total in (0, 1, 7, 24, 25, 100):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_synthetic_generator_matches_oracle_copy():
    import rovr_oracle as O
    from synthetic import masked_frame_batch
    a = masked_frame_batch(3, 64, 48, seed=5)
    b = O.synthetic_localnet_batch(3, 64, 48, seed=5)
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    f, c, t = a
    assert f.shape == (3, 3, 64, 48) and c.shape == (3, 2, 3, 64, 48) and t.shape == (3, 3, 64, 48)
    assert float(f.min()) >= 0 and float(f.max()) <= 1 and (f == 0).any()  # the masked box


def test_grad_arena_tiles_one_storage():
    """_blocks.GradArena (host logic): the gradients of a trunk are views that tile ONE flat fp32 storage completely,
    which is what lets GradientAverager reduce them with a single in-place collective."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                    "reinformcement-optimized-video-reconstruction_b200"))
    from _blocks import GradArena
    P = {"a.weight": torch.zeros(4, 3, 3, 3), "a.bias": torch.zeros(4), "b.weight": torch.zeros(2, 4)}
    arena = GradArena(P)
    views = [arena.take(P["a.weight"]), arena.take(P["a.bias"], zero=True), arena.take(P["b.weight"])]
    assert [tuple(v.shape) for v in views] == [(4, 3, 3, 3), (4,), (2, 4)]
    assert all(v.untyped_storage().data_ptr() == arena.flat.untyped_storage().data_ptr() for v in views)
    assert sum(v.numel() for v in views) * 4 == arena.flat.untyped_storage().nbytes()
    assert float(views[1].abs().sum()) == 0.0
    views[2].fill_(3.0)
    assert float(arena.flat[-8:].sum()) == 24.0
