#!/bin/bash
# ncu --set full captures behind profiles/<tag>_kernel_evidence.md (one representative launch per kernel)
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
run() {   # name, what, kernel regex, skip, count
  timeout 300 python tools/kernel_evidence.py $2 > gpurun_out/ev_plain_$1.log 2>&1 || { echo "plain run $1 failed"; tail -3 gpurun_out/ev_plain_$1.log; return; }
  timeout 600 ncu --set full --clock-control none -k regex:$3 -s $4 -c $5 -f -o gpurun_out/${TAG}_ev_$1 python tools/kernel_evidence.py $2 > gpurun_out/ev_ncu_$1.log 2>&1
  echo "$1 done"
}
WHAT=${2:-all}
sel() { [ "$WHAT" = "all" ] || echo "$WHAT" | grep -qw "$1"; }
sel gemm && run gemm gemm "gemm_tc" 6 3
sel attn && run attn attn "attn_tc" 9 9
sel ln && run ln ln "add_ln" 4 4
sel frontend && run frontend frontend "frontend|cmvn" 3 3
sel adam && run adam adam "adam_kernel" 2 1
sel decode && run decode decode "tree_attn|kv_append" 150 3
sel topk && run topk decode "beam_advance" 50 2
sel decattn && run decattn decode "attn_fwd_smallq|gemm_f32_small" 200 3
true
