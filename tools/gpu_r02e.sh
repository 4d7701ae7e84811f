#!/bin/bash
# round 2: attention forward with O in tensor memory + lazy rescale; both softmax variants.  Stops at the first failure.
set -u
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_attn_tc.py -x -q --timeout 60 2>&1 | tail -5; [ ${PIPESTATUS[0]} -eq 0 ] || { echo "default attention tests failed"; exit 1; }
PKA_ATTN_FAST=1 timeout 150 python -m pytest tests/test_gpu_attn_tc.py -x -q --timeout 60 2>&1 | tail -5; [ ${PIPESTATUS[0]} -eq 0 ] || { echo "fast attention tests failed"; exit 1; }
timeout 200 python tools/bench_attn.py 2>&1 | tail -8 | tee gpurun_out/r02e_attn_default.txt
PKA_ATTN_FAST=1 timeout 200 python tools/bench_attn.py 2>&1 | tail -8 | tee gpurun_out/r02e_attn_fast.txt
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -6 > gpurun_out/r02e_pytest.log; tail -4 gpurun_out/r02e_pytest.log
timeout 200 python tools/profile_decode.py > gpurun_out/r02e_decode_kernels.txt 2>&1; head -30 gpurun_out/r02e_decode_kernels.txt
