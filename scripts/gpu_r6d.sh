#!/bin/bash
TAG=${1:-r6d}
mkdir -p gpurun_out
timeout 300 python scripts/gpu_dec.py 8192 3 > gpurun_out/dec8k_$TAG.log 2>&1; echo "dec8k rc=$?"; tail -6 gpurun_out/dec8k_$TAG.log
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_dec_prepare|k_dec_exec' -s 4 -c 2 -o gpurun_out/prof_dec_$TAG python scripts/gpu_dec.py 8192 1 > gpurun_out/ncu_dec_$TAG.log 2>&1; echo "ncu rc=$?"
