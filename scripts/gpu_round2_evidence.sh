#!/bin/bash
# evidence of the final kernels of round 2: GPU suite, memcheck, bench lines, launch list, ncu captures
TAG=${1:-r9a}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/pytest_gpu_$TAG.log; tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/gpu_sanitize.py > gpurun_out/sanitize_$TAG.log 2>&1; echo "sanitizer rc=$?"; tail -5 gpurun_out/sanitize_$TAG.log
timeout 600 python bench.py --steps 20 --warmup 5 --strong > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_$TAG.err; cut -c1-300 gpurun_out/bench_$TAG.json
timeout 300 python bench.py --impl reference > gpurun_out/bench_${TAG}_reference_arm.json 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches_$TAG.log 2>&1; echo "launch list rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_parse' -s 3 -c 1 -o gpurun_out/prof_parse_$TAG python scripts/gpu_enc.py 8192 1 3 > gpurun_out/ncu_parse_$TAG.log 2>&1; echo "ncu parse rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_entropy' -s 1 -c 1 -o gpurun_out/prof_entropy_$TAG python scripts/gpu_enc.py 8192 1 3 > gpurun_out/ncu_entropy_$TAG.log 2>&1; echo "ncu entropy rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_dec_prepare|k_dec_chains|k_dec_exec' -s 3 -c 3 -o gpurun_out/prof_dec_$TAG python scripts/gpu_dec.py 8192 1 > gpurun_out/ncu_dec_$TAG.log 2>&1; echo "ncu dec rc=$?"
