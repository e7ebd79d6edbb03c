#!/bin/bash
# First-contact diagnostics on a B200: each test group in its own process so that a fault in one
# (a trapped mbarrier timeout kills the CUDA context) cannot hide the others.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/diag_smi.txt 2>&1
for grp in test_pack_and_unpack test_maxpool test_colsum test_tail test_gemm test_conv3x3_fprop test_conv3x3_dgrad test_convT2x2_fprop test_convT2x2_dgrad test_conv3x3_wgrad test_convT2x2_wgrad; do
  echo "=== $grp" | tee -a gpurun_out/diag.log
  timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "$grp" -s 2>&1 | grep -vE "^$|warnings summary|Warning" | tail -25 >> gpurun_out/diag.log
  echo "rc=$?" >> gpurun_out/diag.log
done
echo "=== localnet" >> gpurun_out/diag.log
timeout 600 python -m pytest tests/test_localnet_gpu.py -m gpu -q -s 2>&1 | tail -60 >> gpurun_out/diag.log
tail -150 gpurun_out/diag.log
