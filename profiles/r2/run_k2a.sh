#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_api.py -m gpu -x -q -k "unbinned or golden or mma or batch or single_launch" > gpurun_out/k2a_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/k2a_pytest.log
for i in 1 2; do
timeout 600 python bench.py --steps 30 --warmup 5 --skip-cpu --skip-other > gpurun_out/k2a_bench.json 2> gpurun_out/k2a_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/k2a_bench.json').read().strip().splitlines()[-1])
print("value %.4e ms %.4f e2e %.4e k2_ms %.4f frac %.4f stream %.4f" % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['ms'], d['roofline']['frac'], d['roofline_stream']['frac']))
PY
done
