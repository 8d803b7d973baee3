#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_api.py -m gpu -x -q -k "binned or Binned or beeston or bb" > gpurun_out/k4d_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/k4d_pytest.log
timeout 600 python profiles/r2/diag1.py k4 > gpurun_out/k4d_diag.log 2>&1; echo "auto: $(grep DIAG gpurun_out/k4d_diag.log || tail -5 gpurun_out/k4d_diag.log)"
BI_BINNED_TILE=128 timeout 600 python profiles/r2/diag1.py k4 > gpurun_out/k4d_diag128.log 2>&1; echo "128: $(grep DIAG gpurun_out/k4d_diag128.log || tail -5 gpurun_out/k4d_diag128.log)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_binned|k_canonical" -c 24 --csv --log-file gpurun_out/k4d_p1_launches.csv \
    python profiles/r2/diag1.py k4p1 > /dev/null 2>&1
tail -6 gpurun_out/k4d_p1_launches.csv | cut -d, -f5,14- | cut -c1-150
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_binned|k_canonical" -c 24 --csv --log-file gpurun_out/k4d_p256_launches.csv \
    python profiles/r2/diag1.py k4big > /dev/null 2>&1
tail -6 gpurun_out/k4d_p256_launches.csv | cut -d, -f5,14- | cut -c1-150
