#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/e_pytest.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/e_pytest.log
timeout 600 python profiles/r2/diag1.py c1 c2 k4 k5 > gpurun_out/e_diag.log 2>&1; grep DIAG gpurun_out/e_diag.log || tail -20 gpurun_out/e_diag.log
BI_SMALL=0 BI_SCALAR_FAST=0 timeout 300 python profiles/r2/diag1.py c1 > gpurun_out/e_diag_c1_old.log 2>&1; echo "old path: $(grep DIAG gpurun_out/e_diag_c1_old.log)"
BI_TS_ORDER=0 timeout 300 python profiles/r2/diag1.py k5 > gpurun_out/e_diag_k5_unordered.log 2>&1; echo "k5 toy order: $(grep DIAG gpurun_out/e_diag_k5_unordered.log)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_template_partials -c 1 -f -o gpurun_out/r2_k5_ordered \
    python profiles/r2/diag1.py k5 > gpurun_out/ncu_k5_ordered.log 2>&1
tail -2 gpurun_out/ncu_k5_ordered.log
