#!/bin/bash
# 2 GPUs: peer-Adam kernel after the ILP / grid change, bench nccl vs peer with the DP parity check
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 120 $TR --master-port 29621 tools/check_peer_adam.py --steps 4 --time --out gpurun_out/r02d_peer_adam_n2.json > gpurun_out/r02d_peer_adam.log 2>&1; echo "check_peer_adam exit $?"
grep -E "us_per_step|bit_identical|ok" gpurun_out/r02d_peer_adam.log
for be in nccl peer; do
  PKA_ALLREDUCE=$be timeout 300 $TR --master-port 2963${#be} bench.py --gpus 2 --steps 20 --warmup 5 --no-decode --no-cfg5 \
    > gpurun_out/r02d_bench_n2_${be}.json 2> gpurun_out/r02d_bench_n2_${be}.err; echo "bench $be exit $?"
  tail -2 gpurun_out/r02d_bench_n2_${be}.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02d_bench_n2_${be}.json"))
    print("${be}", d["value"], d["ms_per_step"], d["blocks"]["ms_per_step_all"], d["e2e"]["value"], d.get("dp_parity"), d.get("dp_parity_detail"))
except Exception as e:
    print("no json:", e)
PY
done
