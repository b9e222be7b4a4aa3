#!/bin/bash
TAG=${1:-r5b}
mkdir -p gpurun_out
timeout 300 python scripts/gpu_dec.py 8192 3 > gpurun_out/dec8k_$TAG.log 2>&1; echo "dec8k rc=$?"; tail -7 gpurun_out/dec8k_$TAG.log
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5
