#!/bin/bash
# round 2, final single-GPU pass at HEAD: full GPU suite, smoke(), front-end evidence, the full bench
set -u
mkdir -p gpurun_out
T0=$(date +%s)
timeout 600 python -m pytest tests -m gpu -x -q --timeout 150 --timeout-method=thread 2>&1 | tail -8 > gpurun_out/r02n_pytest_gpu.txt; tail -3 gpurun_out/r02n_pytest_gpu.txt
echo "t=$(( $(date +%s) - T0 ))s pytest"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02n_smoke.txt 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/r02n_smoke.txt
echo "t=$(( $(date +%s) - T0 ))s smoke"
timeout 400 python bench.py > gpurun_out/r02n_bench_n1.json 2> gpurun_out/r02n_bench_n1.err; echo "bench exit $?"
echo "t=$(( $(date +%s) - T0 ))s bench"
timeout 200 python bench.py --impl reference --steps 4 --warmup 3 > gpurun_out/r02n_bench_ref.json 2> gpurun_out/r02n_bench_ref.err; echo "ref exit $?"
echo "t=$(( $(date +%s) - T0 ))s ref"
timeout 200 bash tools/gpu_evidence.sh r02n "frontend"
echo "t=$(( $(date +%s) - T0 ))s evidence"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02n_bench_n1.json'))
print(d['value'], d['ms_per_step'], d['blocks']['ms_per_step_all'], d['e2e']['value'], d['launches_per_step'], d['roofline']['frac'], d['roofline']['family_us'], d['roofline_hbm']['in_step'], d['roofline_hbm']['frac'])
for k in ('decode','cfg5','cfg5_full_attention','cfg5_attention_kernel'):
    print(k, json.dumps(d.get(k))[:900])
print(json.dumps(d['cpu_baseline']))
print(open('gpurun_out/r02n_bench_ref.json').read()[:400])
PY
