#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 200 python tools/bench_attn.py 2>&1 | tail -8 | tee gpurun_out/r02e_attn_default.txt
PKA_ATTN_FAST=1 timeout 200 python tools/bench_attn.py 2>&1 | tail -8 | tee gpurun_out/r02e_attn_fast.txt
timeout 200 python tools/profile_decode.py > gpurun_out/r02f_decode_kernels.txt 2>&1; head -12 gpurun_out/r02f_decode_kernels.txt
timeout 300 python tools/profile_step.py bf16 > gpurun_out/r02f_step_kernels.txt 2>&1; head -14 gpurun_out/r02f_step_kernels.txt
