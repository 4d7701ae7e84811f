#!/bin/bash
# round 2: ncu launch list of the bench command at HEAD + one --set full capture of the dominant kernel inside it
set -u
mkdir -p gpurun_out
timeout 120 python bench.py --steps 2 --warmup 3 --repeats 3 --no-decode --no-cfg5 > gpurun_out/r02q_plain.json 2> gpurun_out/r02q_plain.err; echo "plain exit $?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02q_launches.csv \
  python bench.py --steps 2 --warmup 3 --repeats 3 --no-decode --no-cfg5 > gpurun_out/r02q_ncu_list.log 2>&1; echo "list exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_rows2_kernel -s 60 -c 3 -f -o gpurun_out/r02q_full_gemm_tc_rows2_kernel \
  python bench.py --steps 2 --warmup 3 --repeats 3 --no-decode --no-cfg5 > gpurun_out/r02q_ncu_full.log 2>&1; echo "full exit $?"
