#!/bin/bash
# final validation of the session: GPU tests, smoke, default bench line, reference arm, launch list of the bench command,
# full ncu capture of the K-chunk kernel
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/z_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/z_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/z_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/z_smoke.log
timeout 900 python bench.py > gpurun_out/z_bench.json 2> gpurun_out/z_bench.err; echo "bench rc=$?"; tail -c 400 gpurun_out/z_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/z_ref.json 2> gpurun_out/z_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/z_bench.json').read().strip().splitlines()[-1])
print(json.dumps({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches', 'clocks')}, indent=1))
print(json.dumps({k: d['e2e'][k] for k in ('value', 'ms_per_step')}))
print(json.dumps({k: d['roofline'][k] for k in ('achieved', 'peak', 'frac', 'ms', 'share_of_step')}))
print(json.dumps(d['other_configs'].get('config2_long_contraction'), indent=1))
r = json.loads(open('gpurun_out/z_ref.json').read().strip().splitlines()[-1])
print({k: r[k] for k in ('value', 'ms_per_step')})
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/z_launches.csv python bench.py --steps 2 --warmup 3 --skip-cpu --skip-other > gpurun_out/z_ncu1.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_unbinned_mma_wide -s 3 -c 1 -o gpurun_out/z_wide -f python profiles/r2/wide_probe.py 5 5 50000 4096 > gpurun_out/z_ncu2.log 2>&1; echo "ncu wide rc=$?"
