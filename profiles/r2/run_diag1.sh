#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python profiles/r2/diag1.py > gpurun_out/diag1.log 2>&1 || { tail -30 gpurun_out/diag1.log; exit 1; }
grep DIAG gpurun_out/diag1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_binned_pass -c 2 -f -o gpurun_out/r2_k4_before \
    python profiles/r2/diag1.py k4 > gpurun_out/ncu_k4_before.log 2>&1
tail -2 gpurun_out/ncu_k4_before.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_template_partials -c 1 -f -o gpurun_out/r2_k5_before \
    python profiles/r2/diag1.py k5 > gpurun_out/ncu_k5_before.log 2>&1
tail -2 gpurun_out/ncu_k5_before.log
ls -la gpurun_out
