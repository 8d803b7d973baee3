#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/b2_pytest.log 2>&1; echo "multi pytest rc=$?"; tail -3 gpurun_out/b2_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
    profiles/r2/diag_2gpu.py > gpurun_out/b2_diag.log 2>&1; echo "diag rc=$?"; grep -A30 "Ordered by" gpurun_out/b2_diag.log | head -40; grep DIAG2 gpurun_out/b2_diag.log
