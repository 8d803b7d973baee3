#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/bm2_launches.csv \
    python profiles/r2/diag1.py k5 > gpurun_out/bm2_l.log 2>&1
grep -E "k_bm|k_template|k_point" gpurun_out/bm2_launches.csv | tail -12 | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bm_density -c 1 -f -o gpurun_out/r2_k5c_bm \
    python profiles/r2/diag1.py k5 > gpurun_out/bm2_ncu.log 2>&1
tail -2 gpurun_out/bm2_ncu.log
