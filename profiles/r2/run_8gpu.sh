#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
true
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/g8_bench8.json 2> gpurun_out/g8_bench8.err
echo "bench8 rc=$?"
tail -c 800 gpurun_out/g8_bench8.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/g8_bench8.json').read().strip().splitlines()[-1])
    print(json.dumps({k: d[k] for k in ('value', 'ms_per_step', 'e2e', 'config2_strong_scaling')}, indent=1))
except Exception as e:
    print("parse failed", e)
PY
