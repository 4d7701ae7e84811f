#!/bin/bash
# round 2, call k: front-end swizzle check, in-step kernel breakdown (0.88 -> 0.92 ms regression hunt), attention FAST switch
set -u
mkdir -p gpurun_out
T0=$(date +%s)
timeout 300 python -m pytest tests/test_gpu_ops.py -q -x --timeout 100 --timeout-method=thread 2>&1 | tail -4 > gpurun_out/r02k_pytest.log; tail -2 gpurun_out/r02k_pytest.log
echo "t=$(( $(date +%s) - T0 ))s tests"
timeout 200 python tools/profile_step.py bf16 > gpurun_out/r02k_step_kernels.txt 2>&1; head -34 gpurun_out/r02k_step_kernels.txt | grep -v Warn | tail -30
echo "t=$(( $(date +%s) - T0 ))s profile"
timeout 300 bash tools/gpu_evidence.sh r02k "frontend"
echo "t=$(( $(date +%s) - T0 ))s evidence"
PKA_ATTN_FAST=0 timeout 120 python tools/bench_attn.py 2>&1 | grep -v Warn | tee gpurun_out/r02k_attn_fast0.txt
PKA_ATTN_FAST=1 timeout 120 python tools/bench_attn.py 2>&1 | grep -v Warn | tee gpurun_out/r02k_attn_fast1.txt
PKA_ATTN_FAST=1 timeout 200 python -m pytest tests/test_gpu_attn_tc.py -q -x --timeout 100 --timeout-method=thread 2>&1 | tail -2 | tee gpurun_out/r02k_pytest_fast1.log
echo "t=$(( $(date +%s) - T0 ))s attn"
