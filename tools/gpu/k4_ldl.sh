#!/bin/bash
# the pivot-free K4 attempt: correctness, then timing against its first form and against the pivoted kernels alone
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
rm -f gpurun_out/k4_ldl_time.txt
timeout 600 python tools/check_k4_ldl.py > gpurun_out/k4_ldl_check.txt 2>&1; echo "check rc=$?"; tail -22 gpurun_out/k4_ldl_check.txt | cut -c1-260
timeout 900 python -m pytest tests/test_gpu_solve.py -x -q 2>&1 | tail -5
for n in 64 128 256; do
  for ldl in 1 2 0; do
    echo "FRX_K4_LDL=$ldl" | tee -a gpurun_out/k4_ldl_time.txt
    FRX_K4_LDL=$ldl timeout 300 python tools/time_k4.py --n $n 2>&1 | tail -1 | tee -a gpurun_out/k4_ldl_time.txt
  done
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:frx_k4 --csv --log-file gpurun_out/k4_ldl_launches.csv python tools/time_k4.py --n 128 --steps 1 > gpurun_out/k4_ldl_ncu.log 2>&1
tail -9 gpurun_out/k4_ldl_launches.csv | cut -d, -f5,9,15
