#!/bin/bash
# round 2: cross-KV fusion, fused Adam counters, bucketed padding -- tests, breakdown, full bench
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 240 --timeout-method=thread 2>&1 | tail -15 > gpurun_out/r02c_pytest.log; tail -8 gpurun_out/r02c_pytest.log
timeout 300 python tools/profile_step.py bf16 > gpurun_out/r02c_step_kernels.txt 2>&1; head -45 gpurun_out/r02c_step_kernels.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02c_bench_n1.json 2> gpurun_out/r02c_bench_n1.err; echo "bench exit $?"
tail -3 gpurun_out/r02c_bench_n1.err
cut -c1-1200 gpurun_out/r02c_bench_n1.json
PKA_BENCH_PADDING=set timeout 600 python bench.py --steps 20 --warmup 5 --no-decode --no-cfg5 > gpurun_out/r02c_bench_n1_setpad.json 2>/dev/null; cut -c1-300 gpurun_out/r02c_bench_n1_setpad.json
