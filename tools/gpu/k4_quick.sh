#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python tools/check_k4_ldl.py > gpurun_out/k4_ldl_check.txt 2>&1; echo "check rc=$?"; grep -c '"finite": true' gpurun_out/k4_ldl_check.txt; tail -1 gpurun_out/k4_ldl_check.txt; grep '"n": 256' gpurun_out/k4_ldl_check.txt | head -1 | cut -c1-250
timeout 900 python -m pytest tests/test_gpu_solve.py -x -q 2>&1 | tail -2
for n in 64 128 256; do timeout 300 python tools/time_k4.py --n $n 2>&1 | tail -1 | cut -c1-330; done
