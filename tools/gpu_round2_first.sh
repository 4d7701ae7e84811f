#!/bin/bash
# First GPU session of the next round (2 GPUs): check + time the fused peer-memory optimiser step, then the bench with it.
# usage: gpurun --gpus 2 --timeout 600 -- bash tools/gpu_round2_first.sh
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 120 $TR --master-port 29621 tools/check_peer_adam.py --steps 4 --time --out gpurun_out/r02_peer_adam_n2.json \
  > gpurun_out/peer_adam.log 2>&1; echo "check_peer_adam exit $?" >> gpurun_out/peer_adam.log
tail -40 gpurun_out/peer_adam.log
for be in nccl peer; do
  PKA_ALLREDUCE=$be timeout 200 $TR --master-port 2963${#be} bench.py --gpus 2 --steps 30 --warmup 5 --no-decode --no-cfg5 \
    > gpurun_out/r02_bench_n2_${be}.json 2> gpurun_out/bench_n2_${be}.err; echo "bench $be exit $?"
  cut -c1-400 gpurun_out/r02_bench_n2_${be}.json
done
timeout 200 python -m pytest tests/test_gpu_zz_peer_adam.py -q -x 2>&1 | tail -5
