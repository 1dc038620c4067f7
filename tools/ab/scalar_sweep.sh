#!/bin/bash
# per-launch duration and DRAM bytes of frx_k4_band_scalar / the kernels behind it for a few grid sizes
for n in "$@"; do for W in 1 2 4 8 12; do
  FRX_K4_SCALAR_WARPS=$W ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"band_scalar|bldl" -c 4 --csv --log-file gpurun_out/_sw.csv python tools/time_k4.py --n $n --res-lo 1 --res-hi 6 --steps 1 > /dev/null 2>&1
  python - <<P
import csv
rows=[r for r in csv.reader(l for l in open("gpurun_out/_sw.csv") if l.startswith(chr(34)))]
h=rows[0]; out={}
for r in rows[1:]:
    d=dict(zip(h,r))
    if int(d["ID"])>=2: out.setdefault(d["Kernel Name"][:22],{})[d["Metric Name"][:16]]=d["Metric Value"]
print("n=$n warps/SM=$W", out)
P
done; done
