#!/bin/bash
# ncu --set full on the HBM-bound launches (upconv3 / conv1 fprop, dgrad, wgrad) in isolation
CMD="python scripts/prof_layers.py upconv3 conv1 upconv2"
$CMD > gpurun_out/pl.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"igemm|wgrad" -f -o gpurun_out/prof_membound $CMD > gpurun_out/ncu_mb.log 2>&1
echo rc=$?; tail -2 gpurun_out/ncu_mb.log
ncu -i gpurun_out/prof_membound.ncu-rep --page raw --csv > gpurun_out/prof_membound_raw.csv 2>/dev/null
