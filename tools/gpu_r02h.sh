#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -6 > gpurun_out/r02h_pytest.log; tail -4 gpurun_out/r02h_pytest.log
timeout 300 python tools/profile_step.py bf16 > gpurun_out/r02h_step_kernels.txt 2>&1; head -16 gpurun_out/r02h_step_kernels.txt | tail -14
PKA_SIDE=0 timeout 300 python tools/profile_step.py bf16 2>&1 | head -1
timeout 600 python bench.py --steps 20 --warmup 5 --no-decode --no-cfg5 > gpurun_out/r02h_bench_n1.json 2> gpurun_out/r02h_bench_n1.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/r02h_bench_n1.json')); print(d['value'], d['ms_per_step'], d['blocks']['ms_per_step_all'], d['e2e']['value'], d['launches_per_step'], d['roofline']['frac'], d['roofline_hbm']['in_step'], d['roofline_hbm']['frac'])"
