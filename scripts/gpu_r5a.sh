#!/bin/bash
# r5a: first light of the second-generation staged decoder (k_dec_chains + byte-gather executor)
TAG=${1:-r5a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi_$TAG.log 2>&1
timeout 240 python scripts/gpu_dec.py 1024 3 > gpurun_out/dec1k_$TAG.log 2>&1; echo "dec1k rc=$?"; tail -8 gpurun_out/dec1k_$TAG.log
timeout 300 python scripts/gpu_dec.py 8192 5 > gpurun_out/dec8k_$TAG.log 2>&1; echo "dec8k rc=$?"; tail -8 gpurun_out/dec8k_$TAG.log
timeout 500 python -m pytest tests -x -q -m gpu 2>&1 | tail -12 > gpurun_out/pytest_gpu_$TAG.log; tail -5 gpurun_out/pytest_gpu_$TAG.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_dec_chains|k_dec_exec|k_dec_prepare' -s 6 -c 3 -o gpurun_out/prof_dec_$TAG python scripts/gpu_dec.py 8192 2 > gpurun_out/ncu_dec_$TAG.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_dec_$TAG.log
