#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python profiles/r2/diag1.py k4 > gpurun_out/k4e_diag.log 2>&1; echo "auto: $(grep DIAG gpurun_out/k4e_diag.log || tail -5 gpurun_out/k4e_diag.log)"
BI_BINNED_STORE_LOG2=34 timeout 600 python profiles/r2/diag1.py k4 > gpurun_out/k4e_diag_store.log 2>&1; echo "store always: $(grep DIAG gpurun_out/k4e_diag_store.log || tail -5 gpurun_out/k4e_diag_store.log)"
BI_BINNED_STORE_LOG2=34 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_binned|k_canonical" -c 18 --csv --log-file gpurun_out/k4e_p256_launches.csv \
    python profiles/r2/diag1.py k4big > /dev/null 2>&1
tail -6 gpurun_out/k4e_p256_launches.csv | cut -d, -f5,14- | cut -c1-150
