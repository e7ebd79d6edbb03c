#!/usr/bin/env python
"""Per-launch roofline table of one LocalNet step (B = 24 frames of 256x256) from a bench.py line.

    python scripts/roofline_table.py profiles/r02_bench_final.json > profiles/r02_per_launch_roofline.md

For every launch of the step (the order of `calls_us`, which is the fixed launch order of
local_net._forward_impl / _backward_impl): algorithmic FLOPs, algorithmic HBM bytes (each tensor the
launch must read or write, once; bf16 activations, weights and split-K partials not counted), the two
lower bounds they imply on this GPU (MEASURED_PEAKS.json: burst bf16 tensor peak — the right one for a
kernel timed alone — and HBM copy bandwidth), the measured duration (CUDA events of the un-profiled
run, every kernel alone) and the fraction of the binding bound that the launch reaches.
Reference shapes: rovr/local_net.py:12-39 (channels), :46-72 (the forward graph).
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B, H, W = 24, 256, 256
P1 = B * H * W
P2, P3, P4 = P1 // 4, P1 // 16, P1 // 64
BF = 2      # bytes per bf16


def conv(cin, cout, px, real_cin=None):
    return 2.0 * 9 * (real_cin or cin) * cout * px


def up(cin, cout, px_in):
    return 2.0 * 4 * cin * cout * px_in


# (label, bench name, flops, bytes) in launch order
LAUNCHES = [
    ("pack x, context -> NHWC bf16 (16 ch)", "pack_nchw_to_nhwc", 0, (9 * 4 + 16 * BF) * P1),
    ("conv1 9->64 @256 + ReLU + pool", "conv3x3_fprop_pool2", conv(16, 64, P1, 9), 16 * BF * P1 + 64 * BF * (P1 + P2)),
    ("conv2 64->128 @128 + ReLU + pool", "conv3x3_fprop_pool2", conv(64, 128, P2), 64 * BF * P2 + 128 * BF * (P2 + P3)),
    ("conv3 128->256 @64 + ReLU + pool", "conv3x3_fprop_pool2", conv(128, 256, P3), 128 * BF * P3 + 256 * BF * (P3 + P4)),
    ("conv4 256->512 @32 + ReLU", "conv3x3_fprop", conv(256, 512, P4), (256 + 512) * BF * P4),
    ("upconv1 512->256 32->64 + ReLU", "convT2x2_fprop", up(512, 256, P4), 512 * BF * P4 + 256 * BF * P3),
    ("conv5 512->256 @64 + ReLU", "conv3x3_fprop", conv(512, 256, P3), (512 + 256) * BF * P3),
    ("upconv2 256->128 64->128 + ReLU", "convT2x2_fprop", up(256, 128, P3), 256 * BF * P3 + 128 * BF * P2),
    ("conv6 256->128 @128 + ReLU", "conv3x3_fprop", conv(256, 128, P2), (256 + 128) * BF * P2),
    ("upconv3 128->64 128->256 + ReLU", "convT2x2_fprop", up(128, 64, P2), 128 * BF * P2 + 64 * BF * P1),
    ("conv7 128->64 @256 + ReLU + conv8 + sigmoid + L2", "conv3x3_fprop_tail", conv(128, 64, P1) + 2.0 * 64 * 3 * P1,
     (128 + 64) * BF * P1 + 2 * 3 * 4 * P1),
    ("tail bwd: sigmoid, conv8 dgrad / wgrad, conv7 bias grad", "tail_bwd", 2 * 2.0 * 64 * 3 * P1, (64 + 64) * BF * P1 + 2 * 3 * 4 * P1),
    ("conv7 wgrad", "conv3x3_wgrad", conv(128, 64, P1), (64 + 128) * BF * P1),
    ("conv7 dgrad (+ReLU mask, upconv3 bias grad)", "conv3x3_dgrad", conv(128, 64, P1), (64 + 64 + 128) * BF * P1),
    ("upconv3 wgrad", "convT2x2_wgrad", up(128, 64, P2), 64 * BF * P1 + 128 * BF * P2),
    ("upconv3 dgrad (+mask, conv6 bias grad)", "convT2x2_dgrad", up(128, 64, P2), 64 * BF * P1 + 2 * 128 * BF * P2),
    ("conv6 wgrad", "conv3x3_wgrad", conv(256, 128, P2), (128 + 256) * BF * P2),
    ("conv6 dgrad (+mask, upconv2 bias grad)", "conv3x3_dgrad", conv(256, 128, P2), (128 + 128 + 256) * BF * P2),
    ("upconv2 wgrad", "convT2x2_wgrad", up(256, 128, P3), 128 * BF * P2 + 256 * BF * P3),
    ("upconv2 dgrad (+mask, conv5 bias grad)", "convT2x2_dgrad", up(256, 128, P3), 128 * BF * P2 + 2 * 256 * BF * P3),
    ("conv5 wgrad", "conv3x3_wgrad", conv(512, 256, P3), (256 + 512) * BF * P3),
    ("conv5 dgrad (+mask, upconv1 bias grad)", "conv3x3_dgrad", conv(512, 256, P3), (256 + 256 + 512) * BF * P3),
    ("upconv1 wgrad", "convT2x2_wgrad", up(512, 256, P4), 256 * BF * P3 + 512 * BF * P4),
    ("upconv1 dgrad (+mask, conv4 bias grad)", "convT2x2_dgrad", up(512, 256, P4), 256 * BF * P3 + 2 * 512 * BF * P4),
    ("conv4 wgrad", "conv3x3_wgrad", conv(256, 512, P4), (512 + 256) * BF * P4),
    ("conv4 dgrad", "conv3x3_dgrad", conv(256, 512, P4), (512 + 256) * BF * P4),
    ("pool3 bwd (+skip grad, ReLU mask, conv3 bias grad)", "maxpool_bwd", 0, 3 * 256 * BF * P3 + 256 * BF * P4),
    ("conv3 wgrad", "conv3x3_wgrad", conv(128, 256, P3), (256 + 128) * BF * P3),
    ("conv3 dgrad", "conv3x3_dgrad", conv(128, 256, P3), (256 + 128) * BF * P3),
    ("pool2 bwd (+skip grad, ReLU mask, conv2 bias grad)", "maxpool_bwd", 0, 3 * 128 * BF * P2 + 128 * BF * P3),
    ("conv2 wgrad", "conv3x3_wgrad", conv(64, 128, P2), (128 + 64) * BF * P2),
    ("conv2 dgrad", "conv3x3_dgrad", conv(64, 128, P2), (128 + 64) * BF * P2),
    ("pool1 bwd (+skip grad, ReLU mask, conv1 bias grad)", "maxpool_bwd", 0, 3 * 64 * BF * P1 + 64 * BF * P2),
    ("conv1 wgrad", "conv3x3_wgrad", conv(16, 64, P1, 9), (64 + 16) * BF * P1),
]


def main():
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_bench_final.json")
    with open(path) as f:
        d = json.loads([ln for ln in f if ln.startswith("{")][-1])
    calls = d["calls_us"]
    assert [c[0] for c in calls] == [n for _, n, _, _ in LAUNCHES], "launch order differs from the table"
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    tf = peaks.get("bf16_tflops", 1630.4) * 1e12
    bw = peaks.get("hbm_gbs", 6543.0) * 1e9
    print(f"# Per-launch roofline of one LocalNet step (B = {B}, {H}x{W}) — `{os.path.relpath(path, ROOT)}`\n")
    print(f"Bounds: burst bf16 tensor peak {tf / 1e12:.1f} TFLOP/s, HBM {bw / 1e9:.0f} GB/s (MEASURED_PEAKS.json). "
          "`t_tensor` = algorithmic FLOPs / tensor peak, `t_hbm` = algorithmic bytes / HBM bandwidth, `bound` = the larger "
          "one, `frac` = bound / measured. Measured = CUDA events of the un-profiled run, every kernel alone. "
          "Generated by `scripts/roofline_table.py`.\n")
    print("| # | launch | GFLOP | MB | t_tensor µs | t_hbm µs | bound | measured µs | frac | achieved |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    tot_meas = tot_bound = tot_flops = tot_bytes = 0.0
    for i, ((label, name, flops, nbytes), (_, us)) in enumerate(zip(LAUNCHES, calls), 1):
        tt, th = flops / tf * 1e6, nbytes / bw * 1e6
        bound, kind = (tt, "tensor") if tt >= th else (th, "hbm")
        ach = f"{flops / us / 1e6:.0f} TFLOP/s" if kind == "tensor" else f"{nbytes / us / 1e3:.0f} GB/s"
        print(f"| {i} | {label} | {flops / 1e9:.1f} | {nbytes / 1e6:.0f} | {tt:.1f} | {th:.1f} | {kind} | {us:.1f} | {bound / us:.2f} | {ach} |")
        tot_meas += us
        tot_bound += bound
        tot_flops += flops
        tot_bytes += nbytes
    print(f"| | **sum of the {len(LAUNCHES)} launches** | {tot_flops / 1e9:.0f} | {tot_bytes / 1e6:.0f} | | | | {tot_meas:.0f} | {tot_bound / tot_meas:.2f} | |")
    step_us = d["ms_per_step"] * 1e3
    print(f"\nSum of the per-launch bounds: **{tot_bound:.0f} µs**; sum of the measured launches: {tot_meas:.0f} µs; the step as timed "
          f"({len(calls)} launches + the weight-gradient reductions and the weight re-pack, best launch mode): {step_us:.0f} µs "
          f"— the step reaches **{tot_bound / step_us:.2f} of the roofline** of this launch sequence "
          f"({tot_flops / 1e12 / (step_us * 1e-6):.0f} TFLOP/s model-level).")
    over = [lab for (lab, _, fl, nb), (_, us) in zip(LAUNCHES, calls) if max(fl / tf, nb / bw) * 1e6 > us]
    if over:
        print(f"\n`frac` > 1 ({'; '.join(over)}): the tensor bound is the MEASURED cuBLAS bf16 burst throughput of this pool's "
              "B200s, not the nominal 2250 TFLOP/s — these launches run faster than cuBLAS's own bf16 GEMM.")
    hbm = [(lab, nb / bw * 1e6, us) for (lab, _, fl, nb), (_, us) in zip(LAUNCHES, calls) if nb / bw > fl / tf]
    print(f"\nHBM-bound launches: {len(hbm)} of {len(LAUNCHES)}, {sum(u for _, _, u in hbm):.0f} µs measured against "
          f"{sum(b for _, b, _ in hbm):.0f} µs of bound.")


if __name__ == "__main__":
    main()
