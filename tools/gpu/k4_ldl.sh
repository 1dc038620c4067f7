#!/bin/bash
# first run of the pivot-free K4 attempt: correctness, then timing against the pivoted kernels alone
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python tools/check_k4_ldl.py > gpurun_out/k4_ldl_check.txt 2>&1; echo "check rc=$?"; tail -25 gpurun_out/k4_ldl_check.txt
timeout 900 python -m pytest tests/test_gpu_solve.py -x -q 2>&1 | tail -5
for n in 64 128 256; do
  for ldl in 1 0; do
    FRX_K4_LDL=$ldl timeout 300 python tools/time_k4.py --n $n 2>&1 | tail -1 | tee -a gpurun_out/k4_ldl_time.txt
  done
done
