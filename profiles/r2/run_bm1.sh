#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_template.py -m gpu -x -q > gpurun_out/bm1_pytest.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/bm1_pytest.log
timeout 600 python profiles/r2/diag1.py k5 > gpurun_out/bm1_diag.log 2>&1; grep DIAG gpurun_out/bm1_diag.log || tail -20 gpurun_out/bm1_diag.log
