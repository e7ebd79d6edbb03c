#!/bin/bash
# A/B of one environment knob on the per-call table of a bench step: ab_env.sh VAR v1 v2 ...
VAR=$1; shift
for v in "$@"; do
  env $VAR=$v timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_ab.log 2>/dev/null
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_ab.log') if l.startswith('{')][-1])
print("$VAR=$v value", round(d["value"],1), "ms", round(d["ms_per_step"],3), {k:round(x["ms_per_step"],3) for k,x in d["kernel_classes"].items()})
print(" ".join(f"{n}:{us:.0f}" for n,us in d.get("calls_us",[])))
PY
done
