#!/bin/bash
# usage: gpu_prof_level.sh <tag> <level> <chunks>   -- ncu --set full of the second k_parse launch at that level
TAG=$1; LVL=$2; N=${3:-2048}
mkdir -p gpurun_out
timeout 800 ncu --set full --clock-control none --import-source on -k regex:k_parse -s 1 -c 1 -f -o gpurun_out/prof_parse_L${LVL}_$TAG python scripts/gpu_enc.py $N 1 $LVL > gpurun_out/ncu_parse_L${LVL}_$TAG.log 2>&1
tail -c 400 gpurun_out/ncu_parse_L${LVL}_$TAG.log | tail -3
