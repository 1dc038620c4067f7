#!/bin/bash
# K4 tuning pass on the GPU box: solve tests, then old vs new kernel timings per n and per band-width bin.
out=gpurun_out/k4_compare.jsonl
: > $out
timeout 300 python -m pytest tests/test_gpu_solve.py -x -q -m gpu 2>&1 | tail -15
for n in 64 128 256; do
  for v in 0 2; do
    FRX_K4_VARIANT=$v timeout 120 python tools/time_k4.py --n $n >> $out 2>gpurun_out/k4_err_${n}_$v.log || echo "{\"failed\": \"n=$n variant=$v\"}" >> $out
  done
done
for rng in "1 6" "7 14" "15 22" "23 30"; do
  set -- $rng
  for v in 0 2; do
    FRX_K4_VARIANT=$v timeout 120 python tools/time_k4.py --n 128 --res-lo $1 --res-hi $2 --n-sys 50000 >> $out 2>>gpurun_out/k4_err_bins.log || echo "{\"failed\": \"bin $1-$2 variant=$v\"}" >> $out
  done
done
cat $out
