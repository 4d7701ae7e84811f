#!/bin/bash
# round 2, first validation: whole GPU suite (incl. the new at-size / gate-injected tests), smoke, bench N=1
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -40 > gpurun_out/r02a_pytest.log; tail -15 gpurun_out/r02a_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02a_smoke.log 2>&1; tail -5 gpurun_out/r02a_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench_n1.json 2> gpurun_out/r02a_bench_n1.err; echo "bench exit $?"
tail -3 gpurun_out/r02a_bench_n1.err
cut -c1-1500 gpurun_out/r02a_bench_n1.json
timeout 120 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r02a_bench_ref.json 2>&1; cut -c1-300 gpurun_out/r02a_bench_ref.json
