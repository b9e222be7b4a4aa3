#!/bin/bash
TAG=${1:-x}; N=${2:-8192}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --chunks $N --no-cpu-baseline > gpurun_out/ncu_launches_$TAG.log 2>&1
python - <<PY
import csv,sys
from collections import defaultdict
rows=[r for r in csv.reader(l for l in open('gpurun_out/launches_$TAG.csv') if l.startswith('"'))]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
d=defaultdict(list)
for r in rows[1:]: d[r[ki][:40]].append(float(r[vi].replace(',','')))
for k,v in sorted(d.items(), key=lambda kv:-sum(kv[1])): print(f'{k:42s} n={len(v):3d} avg={sum(v)/len(v)/1e6:9.3f} ms')
PY
