#!/bin/bash
# r6a: streams + async API + bench records on the new decoder
TAG=${1:-r6a}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest_gpu_$TAG.log; tail -6 gpurun_out/pytest_gpu_$TAG.log
timeout 600 python bench.py --strong > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_$TAG.err; cut -c1-1500 gpurun_out/bench_$TAG.json
timeout 300 python bench.py --impl reference > gpurun_out/bench_${TAG}_reference_arm.json 2>/dev/null; cut -c1-400 gpurun_out/bench_${TAG}_reference_arm.json
