#!/bin/bash
TAG=${1:-r6h}
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_dec_chains' -s 2 -c 1 -o gpurun_out/prof_chains_$TAG python scripts/gpu_dec.py 8192 1 > gpurun_out/ncu_chains_$TAG.log 2>&1; echo "ncu rc=$?"
