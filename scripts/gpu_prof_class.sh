#!/bin/bash
# usage: gpu_prof_class.sh <tag> <kernel-regex> <class> <n>
TAG=$1; K=$2; CLS=$3; N=$4
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o gpurun_out/prof_${K}_c${CLS}_$TAG python scripts/gpu_dec_one_class.py $CLS $N > gpurun_out/ncu_${K}_c${CLS}_$TAG.log 2>&1
tail -c 300 gpurun_out/ncu_${K}_c${CLS}_$TAG.log | tail -2
