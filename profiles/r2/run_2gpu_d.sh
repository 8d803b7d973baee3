#!/bin/bash
# multi-rank GPU tests and a short 2-GPU bench after the session's C-ABI changes
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/d2_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/d2_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/d2_bench2.json 2> gpurun_out/d2_bench2.err
echo "bench rc=$?"; tail -c 300 gpurun_out/d2_bench2.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/d2_bench2.json') if l.startswith('{')][-1])
print(json.dumps({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus')}))
print(json.dumps(d['e2e'], indent=1)[:1200])
PY
