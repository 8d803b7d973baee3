#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_template.py tests/test_gpu_toys.py -m gpu -x -q -k "binned or bin_major or toys" > gpurun_out/k4f_pytest.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/k4f_pytest.log
timeout 600 python profiles/r2/diag1.py k4 k5 > gpurun_out/k4f_diag.log 2>&1; echo "fused: $(grep DIAG gpurun_out/k4f_diag.log || tail -20 gpurun_out/k4f_diag.log)"
BI_BINNED_FUSED=0 timeout 600 python profiles/r2/diag1.py k4 > gpurun_out/k4f_diag0.log 2>&1; echo "unfused: $(grep DIAG gpurun_out/k4f_diag0.log || tail -20 gpurun_out/k4f_diag0.log)"
