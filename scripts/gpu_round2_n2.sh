#!/bin/bash
# two ranks over NCCL: the weak-scaling line + the strong-scaling data plane (scatter / gatherv), then the reference arm as the driver launches it
TAG=${1:-r8n2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_n2.json 2> gpurun_out/bench_${TAG}_n2.err; echo "n2 rc=$?"; wc -l gpurun_out/bench_${TAG}_n2.json; head -c 200 gpurun_out/bench_${TAG}_n2.json; echo
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 2 --warmup 1 --impl reference > gpurun_out/bench_${TAG}_n2_reference_arm.json 2>/dev/null; echo "ref rc=$?"; wc -l gpurun_out/bench_${TAG}_n2_reference_arm.json
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
