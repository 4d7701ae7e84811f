#!/bin/bash
# round 2: attention forward with O in tensor memory + lazy rescale; both softmax variants
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_attn_tc.py -x -q --timeout 120 2>&1 | tail -5
PKA_ATTN_FAST=1 timeout 300 python -m pytest tests/test_gpu_attn_tc.py -x -q --timeout 120 2>&1 | tail -5
timeout 200 python tools/bench_attn.py 2>&1 | tail -8 | tee gpurun_out/r02e_attn_default.txt
PKA_ATTN_FAST=1 timeout 200 python tools/bench_attn.py 2>&1 | tail -8 | tee gpurun_out/r02e_attn_fast.txt
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -6 > gpurun_out/r02e_pytest.log; tail -4 gpurun_out/r02e_pytest.log
