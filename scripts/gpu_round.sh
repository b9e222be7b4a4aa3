#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list + full captures of the two big kernels.
# usage: scripts/gpu_round.sh <tag> [skip-tests]
TAG=${1:-r1}
mkdir -p gpurun_out
if [ "$2" != "skip-tests" ]; then
  timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_$TAG.log
fi
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 3000 gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err; tail -c 1500 gpurun_out/bench_ref_$TAG.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --chunks 2048 --no-cpu-baseline > gpurun_out/ncu_launches_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_compress -s 1 -c 1 -f -o gpurun_out/prof_compress_$TAG python bench.py --steps 1 --warmup 1 --chunks 2048 --no-cpu-baseline > gpurun_out/ncu_compress_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_decompress -s 1 -c 1 -f -o gpurun_out/prof_decompress_$TAG python bench.py --steps 1 --warmup 1 --chunks 2048 --no-cpu-baseline > gpurun_out/ncu_decompress_$TAG.log 2>&1
ls -la gpurun_out | tail -20
