#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_ops.py tests/test_gpu_decode.py tests/test_gpu_model.py -x -q --timeout 120 2>&1 | tail -6; [ ${PIPESTATUS[0]} -eq 0 ] || { echo "tests failed"; exit 1; }
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -6 > gpurun_out/r02g_pytest.log; tail -4 gpurun_out/r02g_pytest.log
timeout 300 python tools/profile_step.py bf16 > gpurun_out/r02g_step_kernels.txt 2>&1; head -45 gpurun_out/r02g_step_kernels.txt | tail -42
timeout 200 python tools/profile_decode.py > gpurun_out/r02g_decode_kernels.txt 2>&1; head -8 gpurun_out/r02g_decode_kernels.txt
