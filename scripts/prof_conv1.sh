#!/bin/bash
# ncu --set full + source view of the conv1-shaped forward launch (second iteration) in isolation
CMD="python scripts/prof_layers.py ${1:-conv1}"
$CMD > gpurun_out/pl.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:igemm -s 2 -c 1 -f -o gpurun_out/prof_conv1 $CMD > gpurun_out/ncu_c1.log 2>&1
echo rc=$?; tail -2 gpurun_out/ncu_c1.log
ncu -i gpurun_out/prof_conv1.ncu-rep --page source --csv > gpurun_out/prof_conv1_src.csv 2>/dev/null
ncu -i gpurun_out/prof_conv1.ncu-rep --page raw --csv > gpurun_out/prof_conv1_raw.csv 2>/dev/null
