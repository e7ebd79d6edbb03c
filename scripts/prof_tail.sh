CMD="python bench.py --steps 1 --warmup 3 --profile-run"
$CMD > gpurun_out/plain_t.log 2>&1 && ncu --set full --clock-control none -k regex:"tail_|maxpool_fwd|pack_nchw|wgrad_halo" -s 36 -c 12 -f -o gpurun_out/prof_tail $CMD > gpurun_out/ncu_tail.log 2>&1
echo rc=$?; tail -2 gpurun_out/ncu_tail.log
