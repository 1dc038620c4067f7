#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for n in 64 128 256; do timeout 300 python tools/time_k4.py --n $n 2>&1 | tail -1; done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:frx_k4_band_bldl -s 8 -c 4 -o gpurun_out/r02_k4_band_bldl -f python tools/time_k4.py --n 128 --n-sys 100000 --steps 1 > gpurun_out/k4_ldl_ncu.log 2>&1
echo "full rc=$?"
ls -la gpurun_out/*.ncu-rep
