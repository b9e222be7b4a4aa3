#!/bin/bash
# Final GPU-box visit of the round (r4b): the whole GPU suite exactly as the driver runs it, memcheck over every kernel, launch list.
TAG=${1:-r4b}
mkdir -p gpurun_out
timeout 200 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/pytest_gpu_$TAG.log; tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 180 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/gpu_sanitize.py > gpurun_out/sanitize_$TAG.log 2>&1; echo "sanitizer rc=$?"; tail -8 gpurun_out/sanitize_$TAG.log
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches_$TAG.log 2>&1; tail -2 gpurun_out/ncu_launches_$TAG.log | cut -c1-300
timeout 60 python scripts/gpu_pcie.py > gpurun_out/pcie_$TAG.log 2>&1; cat gpurun_out/pcie_$TAG.log
timeout 60 python scripts/gpu_e2e.py 8192 > gpurun_out/e2e_$TAG.log 2>&1; tail -5 gpurun_out/e2e_$TAG.log
