"""CUDA-graph capture of an arbitrary training / inference step built from the drop-in modules.

The policy networks, the frame-feature extractor and the attention blocks run 65-140 kernels of a
few microseconds each at the reference's batch sizes (b = 20-25): executed eagerly they are bound
by launch latency, not by the GPU. `GraphedFunction` captures `fn(*inputs)` once (forward, loss,
backward, in-place BatchNorm buffer updates, torch's graph-safe CUDA RNG draws for the gumbel noise)
and replays it with new inputs copied into the captured buffers.

    step = GraphedFunction(lambda img, ctx, act: pn1.logprob(img, ctx, act).sum().backward(), (img, ctx, act),
                           modules=[pn1])
    step(img2, ctx2, act2)        # pn1.<param>.grad hold the gradients of this replay

`modules` have their gradients reset (set_to_none) right before the capture, so the captured backward
ASSIGNS fresh gradient tensors instead of accumulating into the ones the warm-up left behind.
`fn` may end with `data_parallel.GradientAverager.average()`: the NCCL all-reduces are then captured
inside the graph (no per-step launch latency); call `close()` before destroying the process group.

Outputs returned by `fn` are static tensors overwritten by every replay (clone to keep them).
"""
import torch


class GraphedFunction:
    def __init__(self, fn, example_inputs, modules=(), warmup=2):
        # fresh leaves: a clone that kept its history would route gradients to the caller's tensor, whose
        # AccumulateGrad node lives on the (non-capturing) stream it was created on
        self.inputs = tuple(t.detach().clone().requires_grad_(t.requires_grad) if torch.is_tensor(t) else t
                            for t in example_inputs)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):     # allocates workspaces, packs weights, warms the allocator
                fn(*self.inputs)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        for m in modules:
            m.zero_grad(set_to_none=True)
        for t in self.inputs:
            if torch.is_tensor(t):
                t.grad = None
        import _native
        n0 = _native.lib.rovr_launch_count()
        self.graph = torch.cuda.CUDAGraph()
        # capture on the SAME stream the warm-up ran on: the per-(device, stream) scratch workspace it sized lives in the
        # ordinary pool and is simply reused — nothing the kernels scribble on is tied to this graph's private pool
        with torch.cuda.graph(self.graph, stream=side):
            self.outputs = fn(*self.inputs)
        self.launches = int(_native.lib.rovr_launch_count() - n0)

    def close(self):
        """Destroy the captured graph (required before `dist.destroy_process_group()` if `fn` issued collectives:
        NCCL does not tear a communicator down while a graph that captured its collectives is alive)."""
        if getattr(self, "graph", None) is not None:
            torch.cuda.synchronize()
            self.graph.reset()
            self.graph = None

    def __call__(self, *inputs):
        for dst, src in zip(self.inputs, inputs):
            if torch.is_tensor(dst) and src is not None:
                dst.data.copy_(src.detach(), non_blocking=True)
        self.graph.replay()
        return self.outputs
