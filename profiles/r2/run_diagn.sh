#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${1:-8}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 \
    profiles/r2/diag_ngpu.py > gpurun_out/diagn.log 2>&1
echo "diag rc=$?"; tail -5 gpurun_out/diagn.log; nproc
