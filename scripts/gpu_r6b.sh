#!/bin/bash
TAG=${1:-r6b}
mkdir -p gpurun_out
timeout 300 python scripts/gpu_enc.py 8192 2 3 > gpurun_out/enc8k_$TAG.log 2>&1; tail -4 gpurun_out/enc8k_$TAG.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_parse' -s 2 -c 1 -o gpurun_out/prof_parse_$TAG python scripts/gpu_enc.py 8192 1 3 > gpurun_out/ncu_parse_$TAG.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_parse_$TAG.log
