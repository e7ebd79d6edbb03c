#!/bin/bash
# Full GPU check: parity tests, then the benchmark. Output in gpurun_out/.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -s 2>&1 | grep -E "^\[|passed|failed|Error|error|assert|bad=[1-9]|hang" | tail -${PYTEST_TAIL:-40} > gpurun_out/pytest.log
tail -12 gpurun_out/pytest.log
python bench.py --steps ${STEPS:-10} --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err
tail -c 1500 gpurun_out/bench.err
python - <<'PY'
import json
try:
    line=[l for l in open('gpurun_out/bench.log') if l.startswith('{')][-1]
    d=json.loads(line)
    print("value", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "eager", round(d["config"].get("eager_ms_per_step_with_per_kernel_events",0),3), "e2e", d["e2e"]["value"], "clocks", d["clocks"])
    print("roofline", {k:(round(v,4) if isinstance(v,float) else v) for k,v in d["roofline"].items() if k in ("achieved","frac","avg_launch_ms")})
    print({k:(round(v["ms_per_step"],3), round(v.get("tflops",0),1)) for k,v in d["kernel_classes"].items()})
    print("model_tflops", round(d["model_tflops"],1), "cpu", d["cpu_baseline"]["value"] if d["cpu_baseline"] else None)
    for i,(n,us) in enumerate(d.get("calls_us",[])): print(f"{i:3d} {n:28s} {us:9.1f}")
except Exception as e:
    print("bench parse failed", e); print(open('gpurun_out/bench.log').read()[-2000:])
PY
