#!/usr/bin/env python
"""bench.py — masked frames/s through the ROVR hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repository's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference arithmetic on host cores

A "step" is one pass of the hot path over one batch of synthetic masked frames: LocalNet U-Net
forward + L2 loss + backward (+ bucketed gradient all-reduce when N > 1), the training step of
rovr/train_local_net_unet.py:102-107,115 restricted to the path BASELINE.json names (no LPIPS, no
optimizer). Workload = BASELINE.json configs[1]: B = 24 frames per GPU, 256x256, synthetic masked
clips (rovr/train_local_net_unet.py:93; rovr/video_ds.py:62-87).

One JSON line is printed by rank 0. Keys beyond the base contract:
  roofline     : aggregate of every launch of the dominant kernel (igemm_kernel: conv fprop, conv
                 dgrad, transposed-conv fprop/dgrad) inside the timed region; achieved = algorithmic
                 FLOPs of those launches / their summed CUDA-event durations; peak = the measured
                 sustained bf16 figure of MEASURED_PEAKS.json (kernel timed inside a long step).
  kernel_classes: the same for the other kernel classes (explains where the step goes).
  cpu_baseline : the oracle port timed on this box's host cores on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "reinformcement-optimized-video-reconstruction_b200")
ORACLE = os.path.join(ROOT, "oracle")
sys.path.insert(0, PKG)

import torch  # noqa: E402

B_PER_GPU, H, W = 24, 256, 256
METRIC = "masked frames/sec (LocalNet fwd+L2+bwd)"
UNIT = "frames/s"

# LocalNet layer table: (kind, cin, cout, output resolution divisor) — rovr/local_net.py:12-39
_LAYERS = [("conv", 9, 64, 1), ("conv", 64, 128, 2), ("conv", 128, 256, 4), ("conv", 256, 512, 8),
           ("up", 512, 256, 8), ("conv", 512, 256, 4), ("up", 256, 128, 4), ("conv", 256, 128, 2),
           ("up", 128, 64, 2), ("conv", 128, 64, 1)]


def algorithmic_macs_per_frame():
    """GEMM MACs per 256x256 frame by kernel class (59.907 GMAC fwd+bwd total,
    including the 1x1 conv8 that runs in the fused tail kernel)."""
    fprop = dgrad = wgrad = 0
    for i, (kind, cin, cout, div) in enumerate(_LAYERS):
        if kind == "conv":
            macs = (H // div) * (W // div) * 9 * cin * cout
        else:  # transposed conv: input resolution H/div, 4 outputs per input pixel
            macs = (H // div) * (W // div) * 4 * cin * cout
        fprop += macs
        wgrad += macs
        if i > 0:  # conv1 has no data gradient (input needs none)
            dgrad += macs
    tail = H * W * 64 * 3
    # conv8 forward runs inside conv7's igemm epilogue (CUDA cores, not counted); the tail class is
    # conv8's data + weight gradient
    return {"igemm": fprop + dgrad, "wgrad": wgrad, "tail": 2 * tail}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"tensor_tflops": p.get("bf16_tflops_sustained", p.get("bf16_tflops")),
                "tensor_tflops_burst": p.get("bf16_tflops"), "hbm_gbs": p.get("hbm_gbs"),
                "source": "measured (MEASURED_PEAKS.json, sustained bf16)"}
    return {"tensor_tflops": 1400.0, "tensor_tflops_burst": 1590.0, "hbm_gbs": 6650.0,
            "source": "fallback (B200_PROFILING.md)"}


# ------------------------------------------------------------------------------------------------
# clocks sampling
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.gpu_index)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        clocks, reasons, mx = [], set(), None
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 8:
                    continue
                try:
                    clocks.append(float(f[1]))
                    mx = float(f[2])
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                      "sw_power_cap"), f[4:8]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if clocks:
            clocks.sort()
            out.update(sm_mhz=clocks[len(clocks) // 2], sm_max_mhz=mx, reasons=sorted(reasons),
                       samples=len(clocks))
        return out


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on host cores
# ------------------------------------------------------------------------------------------------
def cpu_step_rate(steps, warmup, sample_b=None, budget_s=None):
    """Times LocalNet fwd + MSE + bwd of the oracle port (fp32, all host threads) on a bounded
    sample of the workload: `sample_b` frames of 256x256 per step."""
    sys.path.insert(0, ORACLE)
    import rovr_oracle as O
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sd = O.localnet_state_dict(0)
    if sample_b is None:
        # size the sample so one step is about a second on this box
        x, c, t = O.synthetic_localnet_batch(1, H, W, seed=1234)
        O.localnet_step(sd, x, c, t)
        t0 = time.perf_counter()
        O.localnet_step(sd, x, c, t)
        per_frame = time.perf_counter() - t0
        sample_b = max(1, min(B_PER_GPU, int(1.0 / max(per_frame, 1e-3))))
    x, c, t = O.synthetic_localnet_batch(sample_b, H, W, seed=1234)
    for _ in range(warmup):
        O.localnet_step(sd, x, c, t)
    done = 0
    t0 = time.perf_counter()
    for _ in range(steps):
        O.localnet_step(sd, x, c, t)
        done += 1
        if budget_s is not None and time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return {"value": sample_b * done / dt, "unit": UNIT, "cores": torch.get_num_threads(),
            "kind": "port",
            "sample": f"{done} timed step(s) after {warmup} warm-up of the oracle port (PyTorch fp32 CPU, "
                      f"LocalNet fwd+MSE+bwd) on {sample_b} frame(s) of {H}x{W} per step"}, dt / max(done, 1)


def run_reference(args, rank):
    if rank != 0:
        return
    # the stated config: B = 24 frames of 256x256 per step (~2.5 s per step on 16 host threads); the step
    # count is bounded so that the arm ends within a few minutes whatever --steps asks for
    base, ms = cpu_step_rate(min(args.steps, 8), min(args.warmup, 1), sample_b=B_PER_GPU, budget_s=120.0)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "LocalNet U-Net fwd+L2+bwd (configs[1]): B=24 frames, 256x256, synthetic masked "
                                   "clips, random-init weights — on the host cores",
                       "global_batch": B_PER_GPU,
                       "note": "reference is pure PyTorch; its modules cannot travel to the GPU box, so the "
                               "oracle port (pinned to the reference by tests/golden) is what is timed"},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# the CUDA arm
# ------------------------------------------------------------------------------------------------
def synthetic_batch_gpu(dev, seed):
    from synthetic import masked_frame_batch
    return masked_frame_batch(B_PER_GPU, H, W, seed=seed)


def run_cuda(args, rank, local_rank, world):
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import _native
    _native.require_device()
    import ops
    from local_net import LocalNetworkUNetNorm
    from data_parallel import GradientBuckets, broadcast_parameters

    torch.manual_seed(0)
    net = LocalNetworkUNetNorm().to(dev)
    if world > 1:
        broadcast_parameters(net)
        GradientBuckets(net)
    # Host side: frames as a video decoder delivers them — uint8 (rovr/video_ds.py:107-114 reads frames with cv2;
    # the reference then runs ToTensor on the host and ships fp32). Device side: fp32 NCHW in [0, 1] = uint8 / 255,
    # exactly what ToTensor produces; the e2e legs ship the uint8 frames and convert on the GPU (feeder.DeviceFeeder).
    xf, cf, tf = synthetic_batch_gpu(dev, 1234 + rank)
    xh, ch, th = [(v * 255.0).round().to(torch.uint8).pin_memory() for v in (xf, cf, tf)]
    x, c, t = [(v.float() / 255.0).to(dev) for v in (xh, ch, th)]        # host ToTensor == rovr_u8_to_f32 (IEEE division)

    def step_resident(repack=False):
        net.zero_grad(set_to_none=True)
        if repack:      # what an optimizer update causes: every bf16 operand copy is re-packed (one launch)
            net._packed._cache.clear()
        _, loss = net.forward_with_mse(x, c, t)
        loss.backward()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, profile=None):
        barrier()
        ops.set_profile(profile)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ops.set_profile(None)
        ms = e0.elapsed_time(e1)
        if world > 1:
            tms = torch.tensor([ms], device=dev)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            ms = float(tms.item())
        return ms

    # nvidia-smi needs a few hundred ms to start reporting: launch the clock sampler before the
    # warm-up so that it is certainly sampling during the timed regions (warm-up runs the same load)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step_resident()
    # The compute of a step (a fixed sequence of 56 launches) is captured once into a CUDA graph and
    # replayed; when N > 1 the two flat gradient buckets are all-reduced with NCCL right after each
    # replay (inside the timed region). The per-kernel CUDA events behind `roofline` /
    # `kernel_classes` come from an eager pass over the same K steps (events cannot be read inside a
    # graph replay).
    graphed = None
    if not args.no_graph:
        from local_net import GraphedTrainingStep
        # two sets of static inputs / outputs (the step captured twice): the feeder lands batch i+1 in the other set
        # while batch i is computed, so no staging kernel sits between two replays on the compute stream
        graphed = GraphedTrainingStep(net, x, c, t, input_sets=2)
        for _ in range(3):
            graphed()
        for _ in range(2):      # torch.cuda.graph() empties the allocator cache: refill it before timing eager steps
            step_resident()
    prof = []
    n0 = _native.lib.rovr_launch_count()
    esteps = min(args.steps, 20)      # the per-kernel-event pass is bounded (a --steps 1000 run would hold 10^5 events)
    # every kernel is timed ALONE in this pass: the forked weight-gradient stream of the eager mode is switched off, or the
    # events around a data-gradient launch would also span the weight-gradient kernel sharing the SMs with it
    fork_env = os.environ.get("ROVR_WGRAD_STREAM")
    os.environ["ROVR_WGRAD_STREAM"] = "0"
    try:
        ms_eager = timed(step_resident, esteps, profile=prof) * (args.steps / esteps)
    finally:
        if fork_env is None:
            del os.environ["ROVR_WGRAD_STREAM"]
        else:
            os.environ["ROVR_WGRAD_STREAM"] = fork_env
    launches = (_native.lib.rovr_launch_count() - n0) * (args.steps / esteps)
    # Two launch modes of the same kernel sequence are timed over the K steps, both re-packing all 19 weight tensors
    # every step: (1) eager launches through the autograd Function — weight gradients on a forked stream, bucket
    # all-reduces overlapped with backward; (2) ONE CUDA-graph replay per step. `value` is the faster of the two
    # (named in config.launch); both are kept in config.launch_modes_ms_per_step.
    # (at N > 1 only the graph mode is timed when available: measured at N = 2, the eager mode is host-bound there —
    # 3.81 vs 3.28 ms — because every rank also enqueues three NCCL collectives per step from Python)
    n1 = _native.lib.rovr_launch_count()
    time_eager = world == 1 or graphed is None
    ms_eager_plain = timed(lambda: step_resident(repack=True), args.steps) if time_eager else None
    launches_eager = _native.lib.rovr_launch_count() - n1
    ms_graph = timed(graphed, args.steps) if graphed is not None else None
    use_graph = ms_graph is not None and (ms_eager_plain is None or ms_graph <= ms_eager_plain)
    ms_total = ms_graph if use_graph else ms_eager_plain
    launches = graphed.launches_per_step * args.steps if use_graph else launches_eager
    clocks = sampler.stop() if rank == 0 else {}
    frames = B_PER_GPU * world * args.steps
    value = frames / (ms_total * 1e-3)

    e2e_graph = None
    if args.profile_run:
        ms_e2e, e2e_value = float("nan"), None
    else:
        # end to end through the reference-facing nn.Module call, inputs in pinned HOST memory:
        # every step copies its own inputs host -> device (DeviceFeeder: the copy of step i+1 is
        # issued on a side stream before step i is computed) and reads the loss back to the host.
        from feeder import DeviceFeeder, ScalarReadback

        def e2e_loop(steps):
            rb = ScalarReadback(dev, lag=1)
            for xd, cd, td in DeviceFeeder(((xh, ch, th) for _ in range(steps)), dev):
                net.zero_grad(set_to_none=True)
                net._packed._cache.clear()                       # as after optimizer.step(): all operand copies re-packed
                y = net(xd, cd)                                  # the reference-facing call
                loss = torch.nn.functional.mse_loss(y, td)       # rovr/train_local_net_unet.py:107
                loss.backward()
                rb.exchange(loss)            # D2H read of every step's loss, one step behind the enqueue point
            return rb.drain()

        e2e_loop(2)
        ms_e2e = timed(lambda: e2e_loop(args.steps), 1)
        e2e_value = frames / (ms_e2e * 1e-3)
        # the same end-to-end loop through this repository's graphed training-step entry point
        e2e_graph = None
        if graphed is not None:
            def e2e_graph_loop(steps):
                # uint8 frames: H2D + ToTensor (uint8 -> fp32 / 255) on the feeder's stream, straight into the static
                # inputs of the graph set that consumes them; the loss of every step is read back on a side stream
                rb = ScalarReadback(dev, lag=1, side_stream=True)
                feed = DeviceFeeder(((xh, ch, th) for _ in range(steps)), dev, slots=graphed.input_slots)
                for k, _ in enumerate(feed):
                    rb.exchange(graphed.replay(k % 2))
                return rb.drain()
            e2e_graph_loop(2)
            ms_g = timed(lambda: e2e_graph_loop(args.steps), 1)
            e2e_graph = {"value": frames / (ms_g * 1e-3), "unit": UNIT, "ms_per_step": ms_g / args.steps,
                         "api": "step = GraphedTrainingStep(net, ..., input_sets=2); rb = ScalarReadback(lag=1, side_stream=True); "
                                "for k, _ in enumerate(DeviceFeeder(pinned_host_uint8_batches, slots=step.input_slots)): "
                                "rb.exchange(step.replay(k % 2))  # host uint8 frames (as decoded) "
                                "copied H2D every step and converted to fp32/255 on the GPU (ToTensor) into the static inputs, all 19 weight tensors re-packed to bf16 inside the graph every step (as after an "
                                "optimizer update), every step's loss read back on the host"}

    if rank != 0:
        if world > 1:
            from data_parallel import shutdown
            shutdown(graphed)
        return

    # ---- per-kernel-class roofline from the events recorded inside the timed region ----
    peaks = measured_peaks()
    classes = {"igemm": ["rovr_conv3x3_fprop", "rovr_conv3x3_fprop_pool2", "rovr_conv3x3_fprop_tail", "rovr_conv3x3_dgrad", "rovr_convT2x2_fprop", "rovr_convT2x2_dgrad"],
               "wgrad": ["rovr_conv3x3_wgrad", "rovr_convT2x2_wgrad"],
               "tail": ["rovr_tail_fwd", "rovr_tail_bwd"],
               "pool": ["rovr_maxpool_fwd", "rovr_maxpool_bwd"],
               "bias_grad": ["rovr_colsum"],
               "pack": ["rovr_pack_nchw_to_nhwc", "rovr_repack_conv3x3_fprop", "rovr_repack_conv3x3_dgrad",
                        "rovr_repack_convT2x2_fprop", "rovr_repack_convT2x2_dgrad"]}
    dur = {k: 0.0 for k in classes}
    cnt = {k: 0 for k in classes}
    per_name = {}
    for name, a, b in prof:
        ms = a.elapsed_time(b)
        per_name[name] = per_name.get(name, 0.0) + ms
        for k, names in classes.items():
            if name in names:
                dur[k] += ms
                cnt[k] += 1
    # average duration of every call position inside a step (the launch order is fixed)
    per_step = len(prof) // max(esteps, 1)
    call_table = []
    if per_step * esteps == len(prof):
        for i in range(per_step):
            ms = sum(prof[s * per_step + i][1].elapsed_time(prof[s * per_step + i][2]) for s in range(esteps))
            call_table.append([prof[i][0].replace("rovr_", ""), round(ms / esteps * 1e3, 1)])
    macs = algorithmic_macs_per_frame()
    frames_rank0 = B_PER_GPU * esteps
    kc = {}
    for k in classes:
        entry = {"calls_per_step": cnt[k] / esteps, "ms_per_step": dur[k] / esteps}
        if k in macs and dur[k] > 0:
            entry["tflops"] = 2.0 * macs[k] * frames_rank0 / (dur[k] * 1e-3) / 1e12
        kc[k] = entry
    ig_calls = max(cnt["igemm"], 1)
    ig_flops_per_launch = 2.0 * macs["igemm"] * frames_rank0 / ig_calls
    ig_avg_s = dur["igemm"] * 1e-3 / ig_calls
    achieved = ig_flops_per_launch / ig_avg_s / 1e12 if ig_avg_s > 0 else 0.0
    # dram__bytes_read.sum + dram__bytes_write.sum per igemm launch: a STATIC figure taken from the committed
    # `ncu --set full` capture named in profiles/roofline_traffic.json (it cannot be measured in an un-profiled run)
    traffic, traffic_source = None, None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            traffic = tj.get("igemm_dram_bytes_per_launch")
            traffic_source = "static (ncu capture %s)" % tj.get("source", "profiles/")
        except Exception:
            traffic = None
    roofline = {"kernel": "igemm_kernel (tcgen05 implicit GEMM: conv/convT fprop + dgrad)",
                "bound": "tensor", "achieved": achieved, "peak": peaks["tensor_tflops"], "unit": "TFLOP/s",
                "frac": achieved / peaks["tensor_tflops"], "traffic": traffic, "traffic_source": traffic_source,
                "frac_of_burst_peak": achieved / peaks["tensor_tflops_burst"] if peaks.get("tensor_tflops_burst") else None,
                "launches_per_step": cnt["igemm"] / esteps,
                "avg_launch_ms": ig_avg_s * 1e3, "flops_per_launch": ig_flops_per_launch,
                "peak_source": peaks["source"]}
    total_flops_per_frame = 2.0 * (macs["igemm"] + macs["wgrad"] + macs["tail"])
    # the oracle port on this box's host cores at the stated B = 24 (about 2.5 s per step on 16 threads): 1 warm-up + 3 steps
    cpu_base = None if args.profile_run else cpu_step_rate(steps=3, warmup=1, sample_b=B_PER_GPU, budget_s=30.0)[0]

    hb = {"h2d_bytes_per_step": xh.numel() + ch.numel() + th.numel(), "d2h_bytes_per_step": 4}
    e2e_module = None if args.profile_run else dict(
        {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
         "api": "for frame, context, target in DeviceFeeder(pinned_host_uint8_batches): optimizer.zero_grad(); y = "
                "LocalNetworkUNetNorm()(frame, context); loss = F.mse_loss(y, target); loss.backward(); "
                "ScalarReadback.exchange(loss)  # the reference-facing nn.Module call, eager launches (weight gradients on a "
                "forked stream); uint8 frames copied H2D and converted (ToTensor) on the feeder's stream every step, all 19 "
                "bf16 operand copies re-packed every step as after an optimizer update, every loss read back on the host"}, **hb)
    e2e_graphed = dict(e2e_graph, **hb) if e2e_graph is not None else None
    cands = [e for e in (e2e_module, e2e_graphed) if e is not None and e["value"] is not None]
    e2e_best = max(cands, key=lambda e: e["value"]) if cands else None
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "LocalNet U-Net fwd+L2+bwd (configs[1]): B=24 frames/GPU, 256x256, "
                                   "synthetic masked clips, random-init weights",
                       "global_batch": B_PER_GPU * world, "parallelism": f"dp{world}",
                       "launch_modes_ms_per_step": {"eager_forked_wgrad_stream": (ms_eager_plain / args.steps) if ms_eager_plain is not None else None,
                                                    "cuda_graph_replay": (ms_graph / args.steps) if ms_graph is not None else None},
                       "launch": ("eager launches through the autograd Function, weight gradients on a forked stream" +
                                  (", bucketed NCCL all-reduce overlapped with backward" if world > 1 else "")
                                  if not use_graph else "") + (("one CUDA graph per step (GraphedTrainingStep)" +
                                   ((" with the NCCL all-reduces of the 3 gradient buckets captured inside it on a forked "
                                     "stream (decoder and conv4 buckets overlapped with the rest of backward)"
                                     if graphed.allreduce_mode == "captured-overlapped" else
                                     " + NCCL all-reduce of 3 gradient buckets after each replay") if world > 1 else ""))
                                  if use_graph else ""),
                       "eager_ms_per_step_with_per_kernel_events": ms_eager / args.steps,
                       "l2": "each step streams ~2.5 GB of activations/gradients (>> 126 MB L2); no explicit flush",
                       "precision": "bf16 operands + bf16 activation storage, fp32 accumulate, fp32 master weights/grads"},
            "e2e": e2e_best, "e2e_modes": {"module_call_eager": e2e_module, "graphed_step": e2e_graphed},
            "gpu_launches": int(launches), "gpu_launches_per_step": launches / args.steps,
            "clocks": clocks, "roofline": roofline, "kernel_classes": kc,
            "model_tflops": total_flops_per_frame * value / world / 1e12,
            "model_frac_of_peak": total_flops_per_frame * value / world / 1e12 / peaks["tensor_tflops"],
            "cpu_baseline": cpu_base, "calls_us": call_table}
    print(json.dumps(line), flush=True)
    if world > 1:
        from data_parallel import shutdown
        shutdown(graphed)



# ------------------------------------------------------------------------------------------------
# other workloads of the hot path (BASELINE.json configs[2..4] shapes; --workload)
# ------------------------------------------------------------------------------------------------
def _vgg_lpips_gflop_per_frame():
    """LPIPS(net='vgg') at 256x256: VGG16 `features` forward on y_hat AND target (2 x 20.05 GMAC) plus the
    data-gradient chain back to y_hat (20.05 GMAC; VGG is frozen: no weight gradients)."""
    convs = [(3, 64, 1), (64, 64, 1), (64, 128, 2), (128, 128, 2), (128, 256, 4), (256, 256, 4), (256, 256, 4),
             (256, 512, 8), (512, 512, 8), (512, 512, 8), (512, 512, 16), (512, 512, 16), (512, 512, 16)]
    fwd = sum((H // d) * (W // d) * 9 * ci * co for ci, co, d in convs)
    return 2.0 * 3 * fwd / 1e9


WORKLOADS = {
    # name: (metric, unit, description)
    "localnet_lpips": ("masked frames/sec (LocalNet fwd + gamma*L2 + (1-gamma)*LPIPS-VGG + bwd)", "frames/s"),
    "pn2_il": ("policy samples/sec (PolicyNetwork2UNet imitation-learning step fwd+bwd)", "samples/s"),
    "pn1": ("policy samples/sec (PolicyNetwork1UNet logprob fwd+bwd)", "samples/s"),
    "resnet": ("frames/sec (ResnetFeatureExtractor forward + linear fwd/bwd)", "frames/s"),
    "encoder": ("sequences/sec (EncoderBlock E=3072 S=256 fwd+bwd)", "sequences/s"),
    "rovr_step": ("masked frames/sec (full RL step: rollout + PPO, rovr.py restated)", "frames/s"),
}


def build_workload(name, dev, rank, world, args):
    """Returns dict(step=callable (device-resident inputs), e2e=callable(n) -> None (host inputs, n steps),
    units, gflop_per_unit, h2d, d2h, config, close=callable, launches=int|None)."""
    import _native
    from feeder import DeviceFeeder, ScalarReadback
    from graphs import GraphedFunction
    g = torch.Generator().manual_seed(4321 + rank)
    averager = None

    def dp(module):
        nonlocal averager
        if world > 1:
            from data_parallel import GradientAverager, broadcast_parameters
            broadcast_parameters(module)
            averager = GradientAverager(module)

    def reduce_():
        if averager is not None:
            averager.average()

    if name == "localnet_lpips":
        from local_net import GraphedTrainingStep, LocalNetworkUNetNorm
        from lpips_vgg import LPIPS
        from data_parallel import GradientBuckets, broadcast_parameters
        torch.manual_seed(0)
        net = LocalNetworkUNetNorm().to(dev)
        lp = LPIPS(net="vgg").to(dev)
        if world > 1:
            broadcast_parameters(net)
            broadcast_parameters(lp)
            GradientBuckets(net)
        from synthetic import masked_frame_batch
        host = [(t * 255.0).round().to(torch.uint8).pin_memory() for t in masked_frame_batch(B_PER_GPU, H, W, seed=1234 + rank)]
        x, c, t = [(v.float() / 255.0).to(dev) for v in host]      # uint8 frames on the host, ToTensor on the device
        gamma = 0.1 + 0.9 * (0.9993 ** 1000)           # rovr/train_local_net_unet.py:111 at iteration 1000
        step = GraphedTrainingStep(net, x, c, t, lpips_fn=lp, gamma=gamma, input_sets=2)

        def e2e(n):
            rb = ScalarReadback(dev, lag=1, side_stream=True)
            for k, _ in enumerate(DeviceFeeder((tuple(host) for _ in range(n)), dev, slots=step.input_slots)):
                rb.exchange(step.replay(k % 2))
            rb.drain()
        return dict(step=lambda: step(), e2e=e2e, units=B_PER_GPU, gflop_per_unit=119.81 + _vgg_lpips_gflop_per_frame(),
                    h2d=sum(v.numel() for v in host), d2h=4, close=step.close, launches=step.launches_per_step,
                    config={"workload": "LocalNet U-Net training step with the reference's full loss (rovr/train_local_net_unet.py:"
                                        "105-115): B=24 frames/GPU, 256x256, gamma*MSE + (1-gamma)*LPIPS-VGG (random-init VGG16, "
                                        "frozen), forward + backward, one CUDA graph per step",
                            "gamma": gamma, "allreduce": step.allreduce_mode})

    if name == "pn2_il":
        from policy_net_2 import PolicyNetwork2UNet
        torch.manual_seed(0)
        net = PolicyNetwork2UNet().to(dev).train()
        dp(net)
        b = 20                                           # one clip of 20 frames = the reference's batch (imitation_learning.py:83-87)
        enc_h = torch.rand((b, 1, 160, 160), generator=g).pin_memory()
        flat_h = torch.randn((b, 1, 1024), generator=g).pin_memory()
        pos = torch.randint(0, 20, (b, 16, 2), generator=g)
        neg = torch.randint(0, 20, (b, 3, 2), generator=g)
        # multi-hot BCE targets of imitation_learning.py:88-94, built once (they do not depend on the network)
        pos_t = torch.stack([torch.nn.functional.one_hot(pos[:, i], 20).sum(1) for i in range(pos.shape[1])]).float().to(dev)
        neg_t = torch.stack([torch.nn.functional.one_hot(neg[:, i], 20).sum(1) for i in range(neg.shape[1])]).float().to(dev)
        target = torch.arange(20).unsqueeze(1).unsqueeze(1).to(dev)
        bce = torch.nn.functional.binary_cross_entropy_with_logits

        def fn(enc, flat):
            out = net(enc, flat, target, extra=True)
            loss = sum(bce(out, pos_t[i]) * 1.5 for i in range(pos_t.shape[0])) - sum(bce(out, neg_t[i]) for i in range(neg_t.shape[0]))
            loss.backward()
            reduce_()                                    # N > 1: the gradient all-reduces are captured in the graph
            return loss.detach()
        gstep = GraphedFunction(fn, (enc_h.to(dev), flat_h.to(dev)), modules=[net])
        clips = max(1, args.clips)

        def step():
            for _ in range(clips):                       # clips are independent optimizer steps in the reference
                gstep(None, None)

        def e2e(n):
            rb = ScalarReadback(dev, lag=1)
            for encd, flatd in DeviceFeeder(((enc_h, flat_h) for _ in range(n * clips)), dev):
                rb.exchange(gstep(encd, flatd))
            rb.drain()
        return dict(step=step, e2e=e2e, units=b * clips, gflop_per_unit=0.503, h2d=(enc_h.numel() + flat_h.numel()) * 4 * clips,
                    d2h=4 * clips, close=gstep.close, launches=gstep.launches * clips,
                    config={"workload": f"imitation-learning step of rovr/imitation_learning.py:83-100 on PolicyNetwork2UNet: "
                                        f"{clips} clip(s) per GPU per step, each clip = 20 samples of [1,160,160] + [1024] "
                                        "(one forward + BCE loss + backward per clip, per-clip BatchNorm statistics as in the "
                                        "reference; gradients averaged over ranks per clip), emulated-fp32 trunk, one CUDA graph "
                                        "replay per clip", "clips_per_gpu": clips, "precision": net.trunk_precision})

    if name == "pn1":
        from policy_net_1 import PolicyNetwork1UNet
        torch.manual_seed(0)
        net = PolicyNetwork1UNet().to(dev).train()
        dp(net)
        b = 25
        img_h = torch.rand((b, 3, 80, 80), generator=g).pin_memory()
        ctx_h = torch.rand((b, 3, 80, 80), generator=g).pin_memory()
        act = torch.randint(0, 25, (b,), generator=g).to(dev)

        def fn(img, ctx):
            lp = net.logprob(img, ctx, act)
            lp.sum().backward()
            reduce_()
            return lp.detach().sum()
        gstep = GraphedFunction(fn, (img_h.to(dev), ctx_h.to(dev)), modules=[net])

        def step():
            gstep(None, None)

        def e2e(n):
            rb = ScalarReadback(dev, lag=1)
            for i_d, c_d in DeviceFeeder(((img_h, ctx_h) for _ in range(n)), dev):
                rb.exchange(gstep(i_d, c_d))
            rb.drain()
        return dict(step=step, e2e=e2e, units=b, gflop_per_unit=2.960, h2d=(img_h.numel() + ctx_h.numel()) * 4, d2h=4,
                    close=gstep.close, launches=gstep.launches,
                    config={"workload": "PolicyNetwork1UNet.logprob forward + backward, b=25 mosaics of 80x80 (rovr/rovr.py:312 "
                                        "shape), emulated-fp32 trunk, one CUDA graph replay per step",
                            "precision": net.trunk_precision})

    if name == "resnet":
        import warnings
        warnings.filterwarnings("ignore")
        from resnet_extractor import ResnetFeatureExtractor
        torch.manual_seed(0)
        net = ResnetFeatureExtractor(pretrained=False)
        net.resnet.eval()
        for p_ in net.resnet.parameters():
            p_.requires_grad = False
        net = net.to(dev)
        dp(net)
        fr_h = torch.rand((1, 25, 3, 224, 224), generator=g).pin_memory()

        def fn(fr):
            out = net(fr)
            loss = (out ** 2).sum()
            loss.backward()
            reduce_()
            return loss.detach()
        gstep = GraphedFunction(fn, (fr_h.to(dev),), modules=[net])

        def step():
            gstep(None)

        def e2e(n):
            rb = ScalarReadback(dev, lag=1)
            for (fd,) in DeviceFeeder(((fr_h,) for _ in range(n)), dev):
                rb.exchange(gstep(fd))
            rb.drain()
        return dict(step=step, e2e=e2e, units=25, gflop_per_unit=8.174, h2d=fr_h.numel() * 4, d2h=4, close=gstep.close,
                    launches=gstep.launches,
                    config={"workload": "ResnetFeatureExtractor forward of one clip (25 frames, 224x224; frozen eval-mode ResNet-50 "
                                        "trunk as with pretrained=True) + Linear(2048,768) forward/backward + mosaic paste"})

    if name == "encoder":
        from common_layers import EncoderBlock
        torch.manual_seed(0)
        E, S, Bq = 3072, 256, 24
        net = EncoderBlock(E, 8, 0.0).to(dev)
        dp(net)
        x_h = torch.randn((Bq, S, E), generator=g).pin_memory()

        def fn(xx):
            out = net(xx)
            loss = (out ** 2).mean()
            loss.backward()
            reduce_()
            return loss.detach()
        gstep = GraphedFunction(fn, (x_h.to(dev).requires_grad_(True),), modules=[net])

        def step():
            gstep(None)

        def e2e(n):
            rb = ScalarReadback(dev, lag=1)
            for (xd,) in DeviceFeeder(((x_h,) for _ in range(n)), dev):
                rb.exchange(gstep(xd))
            rb.drain()
        return dict(step=step, e2e=e2e, units=Bq, gflop_per_unit=3 * (20.13 + 2.42), h2d=x_h.numel() * 4, d2h=4,
                    close=gstep.close, launches=gstep.launches,
                    config={"workload": "EncoderBlock(hidden 3072, 8 heads) forward + backward on B=24 sequences of 256 tokens "
                                        "(rovr/common_layers.py:94-104 at the token shape of :8-9)"})
    if name == "rovr_step":
        import warnings
        warnings.filterwarnings("ignore")
        from local_net import LocalNetworkUNetNorm
        from lpips_vgg import LPIPS
        from policy_net_2 import PolicyNetwork2UNet
        from rovr_step import ROVRStep
        from video_processor import VideoProcessor
        torch.manual_seed(0)
        actor, critic = PolicyNetwork2UNet().to(dev).train(), PolicyNetwork2UNet(is_critic=True).to(dev).train()
        local, lp, vp = LocalNetworkUNetNorm(freeze=True).to(dev), LPIPS(net="vgg").to(dev), VideoProcessor().to(dev)
        avgs = None
        if world > 1:
            from data_parallel import GradientAverager, broadcast_parameters
            for m in (actor, critic, local, lp, vp):
                broadcast_parameters(m)
            avgs = (GradientAverager(actor), GradientAverager(critic))
        K, S_ = max(1, args.clips), 20
        from synthetic import masked_frame_batch
        fr, _, tg = masked_frame_batch(K * S_, H, W, seed=99 + rank)
        vid_h = (fr.view(K, S_, 3, H, W) * 255).round().to(torch.uint8).pin_memory()
        org_h = (tg.view(K, S_, 3, H, W) * 255).round().to(torch.uint8).pin_memory()
        vid, org = (vid_h.float() / 255).to(dev), (org_h.float() / 255).to(dev)
        rl = ROVRStep(actor, critic, local, lp, vp, averager=avgs, graphed=not args.no_graph)

        def e2e(n):
            rb = ScalarReadback(dev, lag=1)
            for vd, od in DeviceFeeder(((vid_h, org_h) for _ in range(n)), dev):
                losses, _ = rl.train(vd, od)
                rb.exchange(losses[-1][-1][1])
            rb.drain()
        return dict(step=lambda: rl.train(vid, org), e2e=e2e, units=K * S_, gflop_per_unit=0.0,
                    h2d=vid_h.numel() + org_h.numel(), d2h=4, close=lambda: None, launches=None, replayed=rl,
                    config={"workload": f"one RL iteration of rovr/rovr.py:68-337 restated on the drop-ins (rovr_step.ROVRStep): {K} "
                                        "clip(s) per GPU of 20 frames 256x256; rollout = VideoProcessor encode (ResNet-50, 20 frames) + "
                                        "20 time-steps of [PolicyNetwork2UNet actor b=1 per clip, LocalNet forward, LPIPS-VGG reward, "
                                        "re-encode of the reconstructed frame], then PPO on PolicyNetwork2UNet per clip (5 updates of "
                                        "critic fwd+bwd and logprob fwd+bwd at b=20, Adam); " +
                                        ("eager launches" if args.no_graph else "one CUDA graph per time-step (replayed 20x) and two "
                                         "per PPO update") + "; gradients averaged over ranks inside PPO", "clips_per_gpu": K})
    raise ValueError(name)


def run_workload(args, rank, local_rank, world):
    """The generic arm: same timing contract as run_cuda (W >= 3 warm-up steps, K timed steps between barriers +
    synchronize, CUDA events, max over ranks), model-level roofline (algorithmic FLOPs of SURVEY §8d / step time)."""
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import _native
    _native.require_device()
    wl = build_workload(args.workload, dev, rank, world, args)
    metric, unit = WORKLOADS[args.workload]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            tms = torch.tensor([ms], device=dev)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            ms = float(tms.item())
        return ms

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        wl["step"]()
    n0 = _native.lib.rovr_launch_count()
    r0 = wl["replayed"].replayed_launches if wl.get("replayed") is not None else 0
    ms_total = timed(lambda: [wl["step"]() for _ in range(args.steps)])
    # kernels launched by this library in the timed region: counted (eager launches) plus replayed from graphs
    launches_total = (wl["launches"] * args.steps) if wl["launches"] else (_native.lib.rovr_launch_count() - n0)
    if wl.get("replayed") is not None:
        launches_total += wl["replayed"].replayed_launches - r0
    clocks = sampler.stop() if rank == 0 else {}
    wl["e2e"](2)
    ms_e2e = timed(lambda: wl["e2e"](args.steps))
    units = wl["units"] * world * args.steps
    value = units / (ms_total * 1e-3)
    if rank == 0:
        peaks = measured_peaks()
        tflops = wl["gflop_per_unit"] * value / world / 1e3
        line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": dict(wl["config"], parallelism=f"dp{world}",
                               l2="activations of a step exceed the 126 MB L2 only for the LocalNet / ResNet workloads; the "
                                  "policy workloads are launch-bound (working set < L2): no flush applies"),
                "e2e": {"value": units / (ms_e2e * 1e-3), "unit": unit, "ms_per_step": ms_e2e / args.steps,
                        "h2d_bytes_per_step": wl["h2d"], "d2h_bytes_per_step": wl["d2h"],
                        "api": "pinned host inputs -> DeviceFeeder -> graph replay of the module's forward + backward -> "
                               "ScalarReadback of the loss, every step"},
                "gpu_launches": int(launches_total), "gpu_launches_per_step": launches_total / args.steps,
                "clocks": clocks,
                "roofline": ({"kernel": "whole step (model-level): algorithmic GEMM FLOPs of SURVEY §8d / step time",
                              "bound": "tensor", "achieved": tflops, "peak": peaks["tensor_tflops"], "unit": "TFLOP/s",
                              "frac": tflops / peaks["tensor_tflops"], "traffic": None, "peak_source": peaks["source"],
                              "gflop_per_unit": wl["gflop_per_unit"]} if wl["gflop_per_unit"] else None),
                "cpu_baseline": None}
        print(json.dumps(line), flush=True)
    wl["close"]()
    if world > 1:
        from data_parallel import shutdown
        shutdown()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="localnet", choices=["localnet"] + sorted(WORKLOADS),
                    help="localnet (default) = BASELINE.json configs[1], the headline; the others are the remaining "
                         "hot-path rows (SURVEY §8d configs 2-5)")
    ap.add_argument("--clips", type=int, default=1, help="pn2_il: clips per GPU per step (sweep 1/8/64/512)")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of the CUDA-graph replay")
    ap.add_argument("--profile-run", action="store_true",
                    help="for runs under ncu: skip the e2e and cpu_baseline legs (their numbers are null)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run on this node
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 400), os.path.abspath(__file__),
               "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup),
               "--workload", args.workload, "--clips", str(args.clips)] + (["--no-graph"] if args.no_graph else [])
        sys.exit(subprocess.call(cmd))
    if args.workload != "localnet":
        run_workload(args, rank, local_rank, world)
        return
    run_cuda(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
