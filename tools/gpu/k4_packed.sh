#!/bin/bash
# K4 narrow-band bins: one system per warp (FRX_K4_PACKED=0) against several systems per warp (default)
out=gpurun_out/k4_packed.jsonl
: > $out
timeout 300 python -m pytest tests/test_gpu_solve.py -x -q -m gpu 2>&1 | tail -4
for rng in "1 6" "7 14"; do
  set -- $rng
  for pk in 0 3; do
    FRX_K4_PACKED=$pk timeout 120 python tools/time_k4.py --n 128 --res-lo $1 --res-hi $2 --n-sys 50000 >> $out 2>>gpurun_out/k4pk_err.log || echo "{\"failed\": \"bin $1-$2 pk=$pk\"}" >> $out
  done
done
for n in 64 128 256; do
  for pk in 0 3; do
    FRX_K4_PACKED=$pk timeout 120 python tools/time_k4.py --n $n >> $out 2>>gpurun_out/k4pk_err.log || echo "{\"failed\": \"n=$n pk=$pk\"}" >> $out
  done
done
python - <<'PY'
import json
for ln in open('gpurun_out/k4_packed.jsonl'):
    d=json.loads(ln)
    if 'failed' in d: print(d); continue
    print(d['n'], d['res'], 'solve_ms %.3f ms_batch %.3f resid %.1e info %d finite %s'%(d['solve_kernels_ms'], d['ms_per_batch'], d['max_rel_residual'], d['info_nonzero'], d['finite']))
PY
tail -5 gpurun_out/k4pk_err.log
