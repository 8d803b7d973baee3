#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python profiles/profile_driver.py 2 > gpurun_out/p3_k2_plain.log 2>&1 || { tail -5 gpurun_out/p3_k2_plain.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_unbinned_mma -c 1 -f -o gpurun_out/r2_k2_scan_v2 \
    python profiles/profile_driver.py 1 > gpurun_out/p3_k2_ncu.log 2>&1
tail -2 gpurun_out/p3_k2_ncu.log
