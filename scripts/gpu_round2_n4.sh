#!/bin/bash
TAG=${1:-r8n4}; N=${2:-4}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "rc=$?"; wc -l gpurun_out/bench_${TAG}.json; tail -3 gpurun_out/bench_${TAG}.err | cut -c1-300
