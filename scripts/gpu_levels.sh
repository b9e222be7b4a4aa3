#!/bin/bash
# usage: gpu_levels.sh <tag> <chunks> <level>...   -- bench.py at other levels (kernel-side numbers + e2e), plus the CPU arm
TAG=$1; N=$2; shift 2
mkdir -p gpurun_out
for L in "$@"; do
  timeout 600 python bench.py --level $L --chunks $N --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_L$L.json 2> gpurun_out/bench_${TAG}_L$L.err
  timeout 600 python bench.py --impl reference --level $L --chunks $N --steps 2 --warmup 1 > gpurun_out/bench_${TAG}_L${L}_ref.json 2>> gpurun_out/bench_${TAG}_L$L.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_${TAG}_L$L.json').read().strip().splitlines()[-1])
r=json.loads(open('gpurun_out/bench_${TAG}_L${L}_ref.json').read().strip().splitlines()[-1])
print('level', d['config']['level'], 'gpu compress', round(d['compress_gbs'],2), 'decompress', round(d['decompress_gbs'],2), 'e2e', round(d['e2e']['value'],2), 'ratio', round(d['ratio'],3),
      '| cpu compress', round(r['compress_gbs'],2), 'decompress', round(r['decompress_gbs'],2), 'round trip', round(r['value'],2), '|', {k:round(v,1) for k,v in d['kernel_ms'].items() if v > 0.5})
PY
done
