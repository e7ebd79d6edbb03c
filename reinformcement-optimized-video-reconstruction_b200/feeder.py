"""Host -> device input feeder with one batch of look-ahead.

The reference moves every batch with a blocking `.to(device)` at the top of the step
(rovr/train_local_net_unet.py:104,106): 75.5 MB of fp32 frames per 24-frame step, ~1.4 ms over
PCIe that the GPU spends idle. `DeviceFeeder` copies batch i+1 from pinned host memory on a side
stream while batch i is being computed. It owns two sets of static device buffers and alternates
between them, so the steady state allocates nothing (a cudaMalloc / cudaFree in the loop would
serialise the device), and hands batches over with stream-ordered events (no host sync).

The tensors it yields are only valid until the next-but-one batch is requested.
"""
import torch


class DeviceFeeder:
    def __init__(self, batches, device, depth=2, to_float=True, slots=None):
        """batches: iterable of tuples of equally-shaped host tensors (pinned memory recommended).
        slots (optional): `depth` tuples of pre-allocated device tensors to land the batches in instead of the
        feeder's own buffers — e.g. `GraphedTrainingStep(..., input_sets=2).input_slots`, so that a batch arrives
        directly in the static inputs of the graph that will consume it.
        uint8 tensors (frames as decoded by cv2, rovr/video_ds.py:107-114) are shipped as uint8 — a quarter
        of the fp32 bytes — and converted to fp32 / 255 on the device (torchvision's ToTensor) by a kernel
        behind the copy on the feeder's stream, unless to_float=False."""
        self.it = iter(batches)
        self.device = device
        self.depth = depth
        self.to_float = to_float
        self.staging = [None] * depth     # uint8 landing buffers of the slots that convert
        self.stream = torch.cuda.Stream(device=device)
        self.slots = [None] * depth if slots is None else [tuple(sl) for sl in slots]      # static device buffers
        assert len(self.slots) == depth
        self.count = 0
        self._next = None
        self._preload()

    def _preload(self):
        try:
            host = next(self.it)
        except StopIteration:
            self._next = None
            return
        k = self.count % self.depth
        self.count += 1
        conv = [self.to_float and t.dtype == torch.uint8 for t in host]
        if self.slots[k] is None:
            self.slots[k] = tuple(torch.empty(t.shape, dtype=torch.float32 if c else t.dtype, device=self.device)
                                  for t, c in zip(host, conv))
        if self.staging[k] is None:
            self.staging[k] = tuple(torch.empty(t.shape, dtype=torch.uint8, device=self.device) if c else None
                                    for t, c in zip(host, conv))
        # slot k was last read by the step issued `depth` batches ago, whose kernels are already
        # enqueued on the consumer's stream: order the overwrite after them
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            for d, st, h in zip(self.slots[k], self.staging[k], host):
                if st is None:
                    d.copy_(h, non_blocking=True)
                else:
                    import ops
                    st.copy_(h, non_blocking=True)
                    ops.u8_to_f32(st, d)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self._next = (self.slots[k], ev)

    def __iter__(self):
        return self

    def __next__(self):
        if self._next is None:
            raise StopIteration
        dev, ev = self._next
        torch.cuda.current_stream(self.device).wait_event(ev)
        self._preload()          # the next copy overlaps the compute the caller is about to enqueue
        return dev


class ScalarReadback:
    """Device -> host reads of one scalar per step (the loss), `lag` steps behind the enqueue point.

    `float(loss)` right after `backward()` (rovr/train_local_net_unet.py:117 does `loss.item()`)
    blocks the host until the step has finished, so the next step's launches only start once the GPU
    is already idle. `push(loss)` instead enqueues a 4-byte copy into pinned memory behind the step
    and `pop()` returns the OLDEST outstanding value — the host reads step i's loss while step i+1 is
    already queued. Every step's value is still read, in order.
    """

    def __init__(self, device, lag=1, side_stream=False):
        """side_stream=True issues the 4-byte copies on a stream of their own (behind an event of the producing
        stream) so that nothing sits between two graph replays on the compute stream. Only valid if the scalar
        stays untouched until it has been read — e.g. the per-set loss of GraphedTrainingStep(input_sets=2) with
        lag <= 1; a scalar that the very next step overwrites needs the default (same-stream) copy."""
        self.host = torch.empty(lag + 1, dtype=torch.float32).pin_memory()
        self.events = [torch.cuda.Event() for _ in range(lag + 1)]
        self.device = device
        self.side = torch.cuda.Stream(device=device) if side_stream else None
        self.ready = [torch.cuda.Event() for _ in range(lag + 1)]
        self.head = 0      # next slot to write
        self.tail = 0      # oldest unread slot
        self.lag = lag

    def __len__(self):
        return self.head - self.tail

    def push(self, scalar):
        assert len(self) <= self.lag, "pop() before pushing more than lag + 1 values"
        k = self.head % (self.lag + 1)
        if self.side is None:
            self.host[k:k + 1].copy_(scalar.detach().reshape(1), non_blocking=True)
            self.events[k].record(torch.cuda.current_stream(self.device))
        else:
            self.ready[k].record(torch.cuda.current_stream(self.device))
            self.side.wait_event(self.ready[k])
            with torch.cuda.stream(self.side):
                self.host[k:k + 1].copy_(scalar.detach().reshape(1), non_blocking=True)
                self.events[k].record(self.side)
        self.head += 1

    def pop(self):
        k = self.tail % (self.lag + 1)
        self.events[k].synchronize()
        self.tail += 1
        return float(self.host[k])

    def exchange(self, scalar):
        """push(scalar); returns the value that is now `lag` steps old, or None while the pipe fills."""
        self.push(scalar)
        return self.pop() if len(self) > self.lag else None

    def drain(self):
        out = None
        while len(self):
            out = self.pop()
        return out
