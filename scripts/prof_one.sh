#!/bin/bash
# ncu --set full with source for ONE igemm launch: $1 = launches to skip, output gpurun_out/prof_one.ncu-rep
CMD="python bench.py --steps 1 --warmup 3 --profile-run --no-graph"
$CMD > gpurun_out/plain_o.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:igemm_kernel -s ${1:-57} -c ${2:-1} -f -o gpurun_out/prof_one $CMD > gpurun_out/ncu_one.log 2>&1
echo rc=$?; tail -2 gpurun_out/ncu_one.log
