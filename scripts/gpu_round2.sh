#!/bin/bash
# one GPU visit: parity tests, bench line, entropy phases, ordering experiment, other levels
TAG=${1:-x}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_$TAG.log
bash scripts/gpu_fast.sh $TAG 2>&1 | tail -3
timeout 300 python scripts/gpu_phases.py 8192
timeout 300 python scripts/gpu_order_exp.py 8192
bash scripts/gpu_levels.sh $TAG 8192 9 5 1
