#!/bin/bash
# One GPU round trip: build check, parity tests, smoke, bench.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/gpu.txt 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 300 > gpurun_out/pytest_ops.log 2>&1; echo "exit $?" >> gpurun_out/pytest_ops.log
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q --timeout 300 > gpurun_out/pytest_model.log 2>&1; echo "exit $?" >> gpurun_out/pytest_model.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "exit $?" >> gpurun_out/bench.err
tail -5 gpurun_out/smoke.log; tail -15 gpurun_out/pytest_ops.log; tail -15 gpurun_out/pytest_model.log; tail -3 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
