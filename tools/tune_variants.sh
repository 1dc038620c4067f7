#!/bin/bash
# tuning aid: bench the history kernel geometries (FRX_K2_VARIANT) back to back
for v in "$@"; do
  FRX_K2_VARIANT=$v timeout 200 python bench.py --steps 100 --warmup 5 --skip-cpu > gpurun_out/tune_v$v.json 2> gpurun_out/tune_v$v.err
  python -c "
import json
d=json.load(open('gpurun_out/tune_v$v.json')); r=d['roofline']
print('variant', $v, 'ms/step', round(d['ms_per_step'],4), 'main_ms', round(r['kernel_ms'],4), 'TF', round(r['achieved'],2), 'prep_ms', round(r['per_step_kernel_ms']['weights_table'],4), 'e2e_ms', round(d['e2e']['ms_per_step'],3), 'W', d['clocks']['power_w_max'])
" || tail -3 gpurun_out/tune_v$v.err
done
