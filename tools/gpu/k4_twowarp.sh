#!/bin/bash
# K4: one warp per system (FRX_K4_TWO_WARP=0) against two warps per system in the wide bins (default), same batches
out=gpurun_out/k4_twowarp.jsonl
: > $out
timeout 300 python -m pytest tests/test_gpu_solve.py -x -q -m gpu 2>&1 | tail -5
for n in 64 128 256; do
  for tw in 0 12; do
    FRX_K4_TWO_WARP=$tw timeout 120 python tools/time_k4.py --n $n >> $out 2>gpurun_out/k4tw_err_${n}_$tw.log || echo "{\"failed\": \"n=$n tw=$tw\"}" >> $out
  done
done
for rng in "15 22" "23 30"; do
  set -- $rng
  for tw in 0 12; do
    FRX_K4_TWO_WARP=$tw timeout 120 python tools/time_k4.py --n 128 --res-lo $1 --res-hi $2 --n-sys 50000 >> $out 2>>gpurun_out/k4tw_err_bins.log || echo "{\"failed\": \"bin $1-$2 tw=$tw\"}" >> $out
  done
done
cat $out
