#!/bin/bash
# round 2, call i: new decode / CMVN tests, cfg5 kernel breakdown, ncu --set full evidence of the hot kernels, full bench
set -u
mkdir -p gpurun_out
T0=$(date +%s)
timeout 300 python -m pytest tests/test_gpu_decode.py tests/test_gpu_ops.py -q -x --timeout 120 -k "decode or cmvn or frontend or no_cache or noncausal or translate or lattice" 2>&1 | tail -5 > gpurun_out/r02i_pytest_new.log; tail -3 gpurun_out/r02i_pytest_new.log
echo "t=$(( $(date +%s) - T0 ))s tests"
timeout 200 python tools/run_cfg5.py 8 -100 5 profile > gpurun_out/r02i_cfg5_profile.txt 2>&1; tail -c 2500 gpurun_out/r02i_cfg5_profile.txt
echo "t=$(( $(date +%s) - T0 ))s cfg5"
timeout 600 bash tools/gpu_evidence.sh r02 "gemm attn ln frontend decattn"
echo "t=$(( $(date +%s) - T0 ))s evidence"
timeout 600 python bench.py > gpurun_out/r02i_bench_n1.json 2> gpurun_out/r02i_bench_n1.err; echo "bench exit $?"
echo "t=$(( $(date +%s) - T0 ))s bench"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02i_bench_n1.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['launches_per_step'], d['roofline']['frac'], d['roofline_hbm'].get('frac'))
for k in ('decode','cfg5','cfg5_full_attention','cfg5_attention_kernel'):
    print(k, json.dumps(d.get(k))[:700])
PY
