#!/bin/bash
TAG=${1:-r7a}
mkdir -p gpurun_out
timeout 300 python scripts/gpu_enc.py 8192 3 3 > gpurun_out/enc8k_$TAG.log 2>&1; tail -4 gpurun_out/enc8k_$TAG.log
