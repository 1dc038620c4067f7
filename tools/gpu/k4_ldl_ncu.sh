#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:frx_k4 --csv --log-file gpurun_out/k4_ldl_launches.csv python tools/time_k4.py --n 128 --steps 1 > gpurun_out/k4_ldl_ncu.log 2>&1
echo "launch list rc=$?"
tail -12 gpurun_out/k4_ldl_launches.csv | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:frx_k4_band_ldl -s 8 -c 4 -o gpurun_out/r02_k4_band_ldl -f python tools/time_k4.py --n 128 --n-sys 32000 --steps 1 >> gpurun_out/k4_ldl_ncu.log 2>&1
echo "full rc=$?"
ls -la gpurun_out/*.ncu-rep
