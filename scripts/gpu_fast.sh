#!/bin/bash
# quick GPU visit: parity tests + one bench line
TAG=${1:-x}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -12 | tee gpurun_out/pytest_gpu_$TAG.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$TAG.json'))
print({k:d[k] for k in ('value','compress_gbs','decompress_gbs','kernel_ms','ratio')}, 'e2e', d['e2e']['value'], d['clocks'])
PY
tail -3 gpurun_out/bench_$TAG.err
