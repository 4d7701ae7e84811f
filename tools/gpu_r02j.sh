#!/bin/bash
# round 2, call j: LayerNorm / front-end / few-query attention rewrites -- tests first, then ncu evidence and a short bench
set -u
mkdir -p gpurun_out
T0=$(date +%s)
timeout 420 python -m pytest tests/test_gpu_ops.py tests/test_gpu_decode.py tests/test_gpu_tc.py -q -x --timeout 100 --timeout-method=thread 2>&1 | tail -25 > gpurun_out/r02j_pytest.log
tail -6 gpurun_out/r02j_pytest.log
echo "t=$(( $(date +%s) - T0 ))s tests"
if ! grep -q " passed" gpurun_out/r02j_pytest.log || grep -q "failed\|error" gpurun_out/r02j_pytest.log; then
  echo "TESTS NOT GREEN: skipping evidence / bench"; cat gpurun_out/r02j_pytest.log | head -60; exit 1
fi
timeout 420 bash tools/gpu_evidence.sh r02j "ln frontend decattn"
echo "t=$(( $(date +%s) - T0 ))s evidence"
timeout 300 python bench.py --no-cfg5 --steps 20 > gpurun_out/r02j_bench_n1.json 2> gpurun_out/r02j_bench_n1.err; echo "bench exit $?"
echo "t=$(( $(date +%s) - T0 ))s bench"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02j_bench_n1.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['launches_per_step'], d['roofline']['frac'], json.dumps(d['roofline_hbm'])[:900])
print(json.dumps(d.get('decode'))[:1500])
PY
