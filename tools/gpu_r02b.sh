#!/bin/bash
# round 2: batched launches (deferred reductions, operand cache, grouped wgrad, residual epilogue) -- tests, breakdown, bench
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_batched.py -x -q 2>&1 | tail -15
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02b_pytest.log; tail -8 gpurun_out/r02b_pytest.log
timeout 300 python tools/profile_step.py bf16 > gpurun_out/r02b_step_kernels.txt 2>&1; head -60 gpurun_out/r02b_step_kernels.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-decode --no-cfg5 > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err; echo "bench exit $?"
tail -3 gpurun_out/r02b_bench_n1.err
cut -c1-700 gpurun_out/r02b_bench_n1.json
