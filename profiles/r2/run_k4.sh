#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_api.py -m gpu -x -q -k "binned or Binned or beeston or bb" > gpurun_out/k4_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/k4_pytest.log
timeout 600 python profiles/r2/diag1.py k4 > gpurun_out/k4_diag.log 2>&1; grep DIAG gpurun_out/k4_diag.log || tail -20 gpurun_out/k4_diag.log
BI_BINNED_LEGACY=1 timeout 600 python profiles/r2/diag1.py k4 > gpurun_out/k4_diag_legacy.log 2>&1; grep DIAG gpurun_out/k4_diag_legacy.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_binned_tile -c 6 -f -o gpurun_out/r2_k4_tiled \
    python profiles/r2/diag1.py k4 > gpurun_out/ncu_k4_tiled.log 2>&1
tail -2 gpurun_out/ncu_k4_tiled.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ --csv --log-file gpurun_out/c5_launches.csv \
    python profiles/r2/diag1.py c5 > gpurun_out/ncu_c5_launches.log 2>&1
tail -40 gpurun_out/c5_launches.csv | cut -c1-220
