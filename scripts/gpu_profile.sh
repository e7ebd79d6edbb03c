#!/bin/bash
# ncu evidence for the current build: launch list of one bench run + full-set capture of every
# igemm launch of one step. Output in gpurun_out/ (copy summaries into profiles/).
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --profile-run --no-graph"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
# 3 warm-up steps x 19 igemm launches are skipped; the next 19 are one whole step
ncu --set full --clock-control none --import-source on -k regex:igemm_kernel -s 57 -c 19 -f -o gpurun_out/prof_igemm $CMD > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"
tail -3 gpurun_out/ncu_full.log
# the report itself can exceed what gpurun copies back (64 MiB): keep the CSV pages, drop the report
ncu -i gpurun_out/prof_igemm.ncu-rep --page raw --csv > gpurun_out/prof_igemm_raw.csv 2>/dev/null
[ "$(stat -c %s gpurun_out/prof_igemm.ncu-rep)" -gt 40000000 ] && rm -f gpurun_out/prof_igemm.ncu-rep
ls -la gpurun_out
