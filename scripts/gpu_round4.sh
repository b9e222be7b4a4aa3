#!/bin/bash
# One short GPU-box visit (r4): the new -m gpu tests first, then the whole GPU suite, the bench line of both arms, e2e slice sweep.
TAG=${1:-r4a}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_sequences.py tests/test_format.py tests/test_cparams.py -q -m gpu --maxfail=20 2>&1 | tail -40 > gpurun_out/pytest_new_$TAG.log; tail -5 gpurun_out/pytest_new_$TAG.log
timeout 420 python -m pytest tests -q -m gpu --maxfail=10 --deselect tests/test_sequences.py --deselect tests/test_format.py --deselect tests/test_cparams.py 2>&1 | tail -25 > gpurun_out/pytest_gpu_$TAG.log; tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 700 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
timeout 150 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err; tail -c 400 gpurun_out/bench_ref_$TAG.json
timeout 90 python scripts/gpu_e2e.py 8192 > gpurun_out/e2e_$TAG.log 2>&1; tail -6 gpurun_out/e2e_$TAG.log
