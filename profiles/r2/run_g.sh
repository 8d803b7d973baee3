#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/g_pytest.log 2>&1
echo "pytest rc=$?"; tail -16 gpurun_out/g_pytest.log
timeout 300 python profiles/r2/diag1.py c1 > gpurun_out/g_diag_c1.log 2>&1; grep DIAG gpurun_out/g_diag_c1.log || tail -20 gpurun_out/g_diag_c1.log
bash profiles/r2/run_sanitizer.sh
