#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_binned_tile|k_binned_passb" -c 2 -f -o gpurun_out/r2_k4_p1 \
    python profiles/r2/diag1.py k4p1 > gpurun_out/k4l_ncu.log 2>&1
tail -2 gpurun_out/k4l_ncu.log
