#!/bin/bash
# round 2, call m (2 GPUs): gradient-exchange variants of the data-parallel step at HEAD, same box, short runs
set -u
mkdir -p gpurun_out
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 20 --warmup 5 --repeats 5 --no-decode --no-cfg5 --no-parity > gpurun_out/r02m_$tag.json 2> gpurun_out/r02m_$tag.err
  python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/r02m_$tag.json')); print('$tag', round(d['ms_per_step'],4), d['blocks']['ms_per_step_all'], 'e2e', round(d['e2e']['ms_per_step'],4), d['gradient_exchange'])
except Exception as e:
    print('$tag FAILED', e); print(open('gpurun_out/r02m_$tag.err').read()[-600:])
"
}
run prio0_b3 PKA_COMM_PRIO=0
run prio1_b3 PKA_COMM_PRIO=-1
run prio1_b2 PKA_COMM_PRIO=-1 PKA_BUCKETS=2
run prio1_b5 PKA_COMM_PRIO=-1 PKA_BUCKETS=5
run peer PKA_ALLREDUCE=peer
run symm PKA_ALLREDUCE=symm
