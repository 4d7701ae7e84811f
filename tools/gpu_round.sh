#!/bin/bash
# One GPU round trip: smoke, parity tests, bench, kernel breakdown.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/gpu.txt 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "exit $?" >> gpurun_out/bench.err
timeout 300 python tools/profile_step.py > gpurun_out/profile_step.log 2>&1
tail -4 gpurun_out/smoke.log; tail -25 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/bench.log; tail -5 gpurun_out/bench.err; head -60 gpurun_out/profile_step.log | cut -c1-220
