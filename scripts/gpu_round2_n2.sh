#!/bin/bash
# r7c: two ranks over NCCL: the weak-scaling line + the strong-scaling data plane (scatter / gatherv), then the reference arm as the driver launches it
TAG=${1:-r7c}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_${TAG}_n2.json 2> gpurun_out/bench_${TAG}_n2.err; echo "n2 rc=$?"; tail -3 gpurun_out/bench_${TAG}_n2.err; cut -c1-400 gpurun_out/bench_${TAG}_n2.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 2 --warmup 1 --impl reference > gpurun_out/bench_${TAG}_n2_reference_arm.json 2>/dev/null; echo "ref rc=$?"; cut -c1-200 gpurun_out/bench_${TAG}_n2_reference_arm.json
