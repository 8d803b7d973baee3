#!/bin/bash
# round 2, call A (2 GPUs): GPU test suite incl. the multi-rank tests, then a short 2-GPU bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/a_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/a_pytest.log
tail -5 gpurun_out/a_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 2 --steps 10 --warmup 3 --toys 50000 --c5-events 40000000 > gpurun_out/a_bench2.json 2> gpurun_out/a_bench2.err
echo "bench2 rc=$?"
tail -c 1500 gpurun_out/a_bench2.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/a_bench2.json').read().strip().splitlines()[-1])
    print(json.dumps({k: d[k] for k in ('value', 'ms_per_step', 'e2e', 'config2_strong_scaling')}, indent=1))
    oc = d.get('other_configs') or {}
    print(json.dumps(oc.get('config5_event_sharded'), indent=1)[:3000])
    print(json.dumps(oc.get('config4_toys'), indent=1)[:2000])
except Exception as e:
    print("parse failed", e)
PY
