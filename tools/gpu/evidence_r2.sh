#!/bin/bash
# Round-2 evidence on one B200: tests, the default bench line, the ncu launch list of the bench command, and one
# `ncu --set full` capture per shipped kernel (each after its own command has exited 0 without ncu).
set -u
o=gpurun_out
timeout 400 python -m pytest tests -q -m gpu 2>&1 | tail -4
timeout 900 python bench.py --steps 20 --warmup 5 > $o/bench_r2_n1.json 2> $o/bench_r2_n1.err; echo bench rc=$?
timeout 300 python bench.py --steps 3 --warmup 3 --skip-cpu --skip-configs > $o/bench_small.json 2>/dev/null && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/r02_launches_bench_steps3.csv \
    python bench.py --steps 3 --warmup 3 --skip-cpu --skip-configs > $o/ncu_launch.log 2>&1; echo launches rc=$?
NCU="ncu --set full --clock-control none --import-source on"
python tools/prof_history.py 3 > /dev/null && timeout 300 $NCU -k regex:frx_k2_history_main -s 2 -c 1 -o $o/r02_k2_cfg2 python tools/prof_history.py 3 > $o/ncu_k2.log 2>&1; echo k2 rc=$?
timeout 300 $NCU -k regex:frx_k12_prepare -s 2 -c 1 -o $o/r02_k12_prepare python tools/prof_history.py 3 > $o/ncu_k12.log 2>&1; echo k12 rc=$?
timeout 300 $NCU -k regex:frx_k0_alpha -s 2 -c 1 -o $o/r02_k0 python tools/prof_history.py 3 > $o/ncu_k0.log 2>&1; echo k0 rc=$?
python tools/prof_history.py 2 32 > /dev/null && timeout 400 $NCU -k regex:frx_k2_history_main -s 1 -c 1 -o $o/r02_k2_sweep32 python tools/prof_history.py 2 32 > $o/ncu_k2s.log 2>&1; echo k2 sweep rc=$?
python tools/prof_cfg3.py 512 > /dev/null && timeout 600 $NCU -k regex:frx_k2_history_main -s 2 -c 1 -o $o/r02_k3_cfg3_b512 python tools/prof_cfg3.py 512 > $o/ncu_k3.log 2>&1; echo k3 rc=$?
python tools/prof_stencils.py > /dev/null && timeout 300 $NCU -k regex:frx_k5 -s 2 -c 3 -o $o/r02_k5 python tools/prof_stencils.py > $o/ncu_k5.log 2>&1; echo k5 rc=$?
python tools/time_k4.py --n 128 --n-sys 20000 --steps 1 > $o/k4_time.json && timeout 400 $NCU -k regex:frx_k4_band_dmma -c 4 -o $o/r02_k4_band_dmma python tools/time_k4.py --n 128 --n-sys 20000 --steps 1 > $o/ncu_k4.log 2>&1; echo k4 rc=$?
ls -la $o/*.ncu-rep | tail -12
