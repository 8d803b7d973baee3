#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_api.py tests/test_gpu_toys.py -m gpu -x -q -k "binned" > gpurun_out/k4h_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/k4h_pytest.log
timeout 600 python profiles/r2/diag_k4e2e.py > gpurun_out/k4h_diag.log 2>&1; grep DIAG gpurun_out/k4h_diag.log || tail -20 gpurun_out/k4h_diag.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/k4h_p1.csv \
    python profiles/r2/diag1.py k4p1 > gpurun_out/k4h_l.log 2>&1
tail -7 gpurun_out/k4h_p1.csv | cut -d'"' -f10,16,30 | cut -c1-100
