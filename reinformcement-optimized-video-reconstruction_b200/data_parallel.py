"""Data parallelism for the ROVR hot path: one process per GPU, parameters replicated, the video
batch sharded by frame (LocalNet) or by whole clip (policy nets), gradients averaged with one
all-reduce per bucket, overlapped with the rest of backward.

The reference has no distributed code at all (`parallel_and_device()` is just
`.to(device)`, rovr/train_local_net_unet.py:73-75), so this is new work specified by §8e:

  * LocalNet's backward produces its gradients decoder-first. `local_net._LocalNetFunction`
    writes them into three flat fp32 buckets (decoder: conv8..upconv1 = 2 237 507 elements, ready
    first; conv4 = 1 180 160, ready one kernel later; conv3..conv1 = 374 272) and calls
    `_bucket_ready(i, flat)` the moment a bucket is complete. `GradientBuckets` then launches the
    all-reduce of that bucket on a side stream (NCCL over NVLink/NVSwitch) while the rest of
    backward is still running — only the 1.5 MB tail bucket is exposed at the end — and makes the
    main stream wait for all of them before autograd accumulates into `.grad`.
  * No other collective exists on the path (no all-gather / all-to-all): frames are independent.
  * Policy nets use train-mode BatchNorm; they are sharded by whole clip so each replica sees
    exactly the per-clip batch of the reference (§8e caveat) and only gradients are reduced.
"""
import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """Contiguous [lo, hi) slice of `total` items owned by `rank` (remainder to the low ranks)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class GradientBuckets:
    """Bucketed, overlapped gradient averaging for a module that exposes the
    `_grad_bucket_hook` / `_grad_bucket_wait` hook points (LocalNetworkUNetNorm)."""

    def __init__(self, module, process_group=None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.group = process_group
        self.world = dist.get_world_size(process_group)
        self.backend = dist.get_backend(process_group)
        self.on_gpu = self.backend == "nccl"
        self.stream = torch.cuda.Stream() if self.on_gpu else None   # on the process's current device
        self.pending = []
        self.launched = 0
        module._grad_bucket_hook = self.ready
        module._grad_bucket_wait = self.wait
        module._grad_bucket_reduce = self.reduce_now

    def ready(self, index, flat):
        if self.world == 1:
            return
        self.launched += 1
        if self.on_gpu:
            main = torch.cuda.current_stream()
            self.stream.wait_stream(main)          # the bucket's producers have been enqueued
            if not torch.cuda.is_current_stream_capturing():
                flat.record_stream(self.stream)    # (a captured graph owns its buffers for its lifetime)
            with torch.cuda.stream(self.stream):
                dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group)
            self.pending.append(flat)
        else:
            # gloo (CPU tests of the host logic): SUM then scale
            work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self.pending.append((work, flat))

    def reduce_now(self, flats):
        """Average the given flat gradient buckets on the current stream (used after a CUDA-graph
        replay of the compute, where the per-bucket hooks do not fire)."""
        if self.world == 1:
            return
        for flat in flats:
            if self.on_gpu:
                dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group)
            else:
                dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
                flat.div_(self.world)

    def wait(self):
        if self.world == 1:
            return
        if self.on_gpu:
            torch.cuda.current_stream().wait_stream(self.stream)
        else:
            for work, flat in self.pending:
                work.wait()
                flat.div_(self.world)
        self.pending = []


class GradientAverager:
    """Gradient averaging for ANY module after `loss.backward()` (policy networks, the frame-feature
    projection, attention blocks): the IL / PPO steps of rovr/imitation_learning.py:83-100 and
    rovr/rovr.py:299-334 sharded by whole clip (train-mode BatchNorm and the batch-dim
    standardisation keep per-replica statistics = the reference's per-clip batch; only gradients are
    reduced).

    The trunk Functions write all their parameter gradients into one flat arena (_blocks.GradArena), so
    every gradient that shares a storage with others is reduced by ONE collective over that storage; the
    remaining (head) gradients are flattened together into one more. Call `average()` after backward."""

    def __init__(self, module, process_group=None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.module = module
        self.group = process_group
        self.world = dist.get_world_size(process_group)
        self.on_gpu = dist.get_backend(process_group) == "nccl"
        self.collectives = 0

    def _allreduce(self, flat):
        self.collectives += 1
        if self.on_gpu:
            dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group)
        else:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            flat.div_(self.world)

    def average(self):
        if self.world == 1:
            return
        groups, loose = {}, []
        for p in self.module.parameters():
            g = p.grad
            if g is None:
                continue
            st = g.untyped_storage()
            if st.nbytes() > g.numel() * g.element_size() and g.dtype == torch.float32 and g.is_contiguous():
                groups.setdefault(st.data_ptr(), [st, []])[1].append(g)
            else:
                loose.append(g)
        for st, grads in groups.values():
            # an arena = a storage that these gradients tile completely (nothing else lives in it): one collective
            # over the whole storage, in place. Anything else (a gradient that is a view of some larger tensor)
            # goes the flatten / copy-back way so that no foreign data is ever averaged.
            if sum(g.numel() * 4 for g in grads) == st.nbytes():
                self._allreduce(torch.empty(0, dtype=torch.float32, device=grads[0].device).set_(st))
            else:
                loose.extend(grads)
        if loose:
            flat = torch.cat([g.reshape(-1) for g in loose])
            self._allreduce(flat)
            off = 0
            for g in loose:
                g.copy_(flat[off:off + g.numel()].view_as(g))
                off += g.numel()


def shutdown(*graphed_steps, grace_s=30.0):
    """Orderly end of a data-parallel process: destroy the CUDA graphs that captured collectives, then the
    process group. NCCL's communicator teardown can block if anything still references captured
    collectives, so a daemon timer ends the process (exit code 0, all results are already written by
    then) if the teardown has not returned after `grace_s` seconds."""
    import os
    import sys
    import threading
    for g in graphed_steps:
        if g is not None:
            g.close()
    if not dist.is_initialized():
        return
    sys.stdout.flush()
    sys.stderr.flush()
    timer = threading.Timer(grace_s, lambda: os._exit(0))
    timer.daemon = True
    timer.start()
    try:
        torch.cuda.synchronize() if torch.cuda.is_available() else None
        dist.barrier()
        dist.destroy_process_group()
    finally:
        timer.cancel()


def broadcast_parameters(module, src=0, process_group=None):
    """Make every replica start from rank `src`'s parameters and buffers. The broadcast writes the
    parameter itself (under no_grad), which bumps its version counter, so the cached bf16 operand
    copies keyed on (data_ptr, _version) are re-packed on the next forward."""
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t, src=src, group=process_group)
    for m in module.modules():
        cache = getattr(m, "_packed", None)
        if cache is not None and hasattr(cache, "_cache"):
            cache._cache.clear()
