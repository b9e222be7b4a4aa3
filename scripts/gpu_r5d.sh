#!/bin/bash
TAG=${1:-r5d}
mkdir -p gpurun_out
timeout 300 python scripts/gpu_dec.py 8192 3 > gpurun_out/dec8k_$TAG.log 2>&1; echo "dec8k rc=$?"; tail -7 gpurun_out/dec8k_$TAG.log
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_dec_chains|k_dec_exec' -s 4 -c 2 -o gpurun_out/prof_dec_$TAG python scripts/gpu_dec.py 8192 1 > gpurun_out/ncu_dec_$TAG.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_dec_$TAG.log
