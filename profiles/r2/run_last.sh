#!/bin/bash
# last validation of the round: GPU tests, smoke, default bench line
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/l_pytest.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/l_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/l_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/l_smoke.log
timeout 600 python bench.py > gpurun_out/l_bench.json 2> gpurun_out/l_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/l_bench.json').read().strip().splitlines()[-1])
print(json.dumps({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}))
print(json.dumps({k: d['e2e'][k] for k in ('value', 'ms_per_step')}))
print(json.dumps({k: d['roofline'][k] for k in ('achieved', 'peak', 'frac', 'ms')}))
oc = d['other_configs']
print(oc['config3_binned_bb']['single_point']['device_ms'], oc['config3_binned_bb']['scan']['device_ms'])
print(oc['config2_long_contraction']['evaluation_ms'], oc['config2_long_contraction']['roofline']['frac'])
PY
