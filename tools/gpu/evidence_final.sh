#!/bin/bash
# final single-GPU evidence of round 2: smoke, GPU suite, bench line, ncu launch list of one bench step, full captures of K4 (two paths) and K5
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2_reference.json 2>> gpurun_out/bench_r2_n1.err; echo "reference arm rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:frx_k --csv --log-file gpurun_out/r02_k4_two_path_launches.csv python tools/time_k4.py --n 128 --steps 1 > gpurun_out/k4_ldl_ncu.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"frx_k4_band|frx_k5" -s 11 -c 11 -o gpurun_out/r02_k4_two_path -f python tools/time_k4.py --n 128 --n-sys 100000 --steps 1 >> gpurun_out/k4_ldl_ncu.log 2>&1
echo "full rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -3
