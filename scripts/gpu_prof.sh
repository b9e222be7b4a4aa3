#!/bin/bash
# usage: gpu_prof.sh <tag> <kernel-regex> [chunks]
TAG=$1; K=$2; N=${3:-2048}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o gpurun_out/prof_${K}_$TAG python bench.py --steps 1 --warmup 1 --chunks $N --no-cpu-baseline > gpurun_out/ncu_${K}_$TAG.log 2>&1
tail -c 300 gpurun_out/ncu_${K}_$TAG.log | tail -2
