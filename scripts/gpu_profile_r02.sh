#!/bin/bash
# ncu evidence for the round-2 build: (1) launch list of one eager bench run, (2) full-set capture of the 19 igemm
# launches of one LocalNet step (+ the wgrad kernels), (3) the sustained bench line. Output in gpurun_out/.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --profile-run --no-graph"
timeout 300 $CMD > gpurun_out/r2_plain.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu_launch.log 2>&1
echo "launch list rc=$?"
# 3 warm-up steps x 19 igemm launches are skipped; the next 19 are one whole step
timeout 900 ncu --set full --clock-control none --import-source on -k regex:igemm_kernel -s 57 -c 19 -f -o gpurun_out/r2_prof_igemm $CMD > gpurun_out/r2_ncu_full.log 2>&1
echo "full rc=$?"
ncu -i gpurun_out/r2_prof_igemm.ncu-rep --page raw --csv > gpurun_out/r2_prof_igemm_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none -k regex:wgrad -s 30 -c 10 -f -o gpurun_out/r2_prof_wgrad $CMD > gpurun_out/r2_ncu_wgrad.log 2>&1
ncu -i gpurun_out/r2_prof_wgrad.ncu-rep --page raw --csv > gpurun_out/r2_prof_wgrad_raw.csv 2>/dev/null
rm -f gpurun_out/r2_prof_igemm.ncu-rep gpurun_out/r2_prof_wgrad.ncu-rep
timeout 300 python bench.py --steps 1000 --warmup 10 > gpurun_out/r2_bench_sustained.json 2> gpurun_out/r2_bench_sustained.err
echo "sustained rc=$?"
ls -la gpurun_out | tail -8
