#!/bin/bash
# Weak-scaling lines of every bench workload on N GPUs of one node (N = $1): one JSON line per workload in
# gpurun_out/r2_scale_${N}gpu.jsonl. Each run is bounded by `timeout`.
N=${1:-2}
OUT=gpurun_out/r2_scale_${N}gpu.jsonl
: > $OUT
run() {
  timeout 240 python bench.py --gpus $N "$@" >> $OUT 2>> gpurun_out/r2_scale_${N}gpu.err
  echo "rc=$? $*" >> gpurun_out/r2_scale_${N}gpu.err
}
run --steps 20 --warmup 3
run --workload localnet_lpips --steps 10 --warmup 3
run --workload pn2_il --clips 8 --steps 5 --warmup 3
run --workload pn1 --steps 20 --warmup 3
run --workload resnet --steps 20 --warmup 3
run --workload encoder --steps 10 --warmup 3
run --workload rovr_step --clips 2 --steps 2 --warmup 3
wc -l $OUT
