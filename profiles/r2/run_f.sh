#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/f_pytest.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/f_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/f_smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/f_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f_ref.json 2> gpurun_out/f_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/f_bench.json').read().strip().splitlines()[-1])
print(json.dumps({k: d[k] for k in ('value', 'ms_per_step', 'e2e', 'roofline', 'cpu_baseline', 'clocks')}, indent=1)[:3500])
oc = d['other_configs']
print(json.dumps(oc['config1_gaussian'], indent=1))
print(json.dumps(oc['config3_binned_bb'], indent=1)[:2500])
print(json.dumps(oc['config4_toys'], indent=1)[:1500])
r = json.loads(open('gpurun_out/f_ref.json').read().strip().splitlines()[-1])
print({k: r[k] for k in ('value', 'ms_per_step', 'cpu_baseline')})
PY
