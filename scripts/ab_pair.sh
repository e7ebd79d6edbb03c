#!/bin/bash
# A/B of CTA-pair (cta_group::2) launches on the per-call table of one bench step
for v in 0 1; do
  ROVR_PAIR=$v timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_pair$v.log 2>/dev/null
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_pair$v.log') if l.startswith('{')][-1])
print("pair=$v value", round(d["value"],1), "ms", round(d["ms_per_step"],3), {k:round(x["ms_per_step"],3) for k,x in d["kernel_classes"].items()})
print(" ".join(f"{n}:{us:.0f}" for n,us in d.get("calls_us",[])))
PY
done
