#!/bin/bash
mkdir -p gpurun_out
timeout 500 python scripts/gpu_stress_dec.py 60 > gpurun_out/stress_dec.log 2>&1; echo "stress rc=$?"; tail -4 gpurun_out/stress_dec.log
timeout 300 python scripts/gpu_dec.py 8192 12 > gpurun_out/dec8k_repeat.log 2>&1; echo "repeat rc=$?"; grep -c "rep " gpurun_out/dec8k_repeat.log; tail -3 gpurun_out/dec8k_repeat.log | cut -c1-200
