#!/bin/bash
# validation after the retirement of k_unbinned_grouped: GPU tests, smoke, default bench line
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/v_pytest.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/v_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/v_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/v_smoke.log
timeout 900 python bench.py > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/v_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/v_bench.json').read().strip().splitlines()[-1])
print(json.dumps({k: d[k] for k in ('value', 'ms_per_step', 'e2e', 'roofline', 'clocks')}, indent=1)[:3000])
PY
