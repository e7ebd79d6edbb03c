#!/bin/bash
# A/B of the epilogue staging depth on the per-call table of one bench step
for v in 1 2; do
  ROVR_STG_BUFS=$v python bench.py --steps 10 --warmup 3 > gpurun_out/bench_stg$v.log 2>/dev/null
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_stg$v.log') if l.startswith('{')][-1])
print("stg_bufs=$v value", round(d["value"],1), "ms", round(d["ms_per_step"],3), {k:round(x["ms_per_step"],3) for k,x in d["kernel_classes"].items()})
print(" ".join(f"{n}:{us:.0f}" for n,us in d.get("calls_us",[])))
PY
done
