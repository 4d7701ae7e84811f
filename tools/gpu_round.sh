#!/bin/bash
# One GPU round trip: smoke, parity tests, bench, ncu launch list + one full capture.  Everything lands in gpurun_out/.
# usage: tools/gpu_round.sh [tag] [kernel-regex-for-full-capture]
set -u
TAG=${1:-r01}
KRE=${2:-gemm_tc_kernel}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/gpu.txt 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/bench.err; echo "exit $?" >> gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/bench.err
# ncu launch list of the same command (graph kernel nodes are profiled individually); numbers under ncu are not bench values
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-decode --no-cfg5 > gpurun_out/ncu_list.log 2>&1
if [ "$KRE" != "skip" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:${KRE} -s 40 -c 3 -f -o gpurun_out/${TAG}_full_${KRE} \
  python bench.py --steps 2 --warmup 3 --no-decode --no-cfg5 > gpurun_out/ncu_full.log 2>&1
fi
tail -4 gpurun_out/smoke.log; tail -15 gpurun_out/pytest_gpu.log; cut -c1-1500 gpurun_out/${TAG}_bench_n1.json; tail -5 gpurun_out/bench.err
