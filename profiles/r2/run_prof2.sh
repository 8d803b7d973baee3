#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python profiles/r2/prof_k5b.py > gpurun_out/p2_k5b_plain.log 2>&1 || { tail -5 gpurun_out/p2_k5b_plain.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mixture_partials_mma -c 1 -f -o gpurun_out/r2_k5b_p11_v2 \
    python profiles/r2/prof_k5b.py > gpurun_out/p2_k5b_ncu.log 2>&1
tail -2 gpurun_out/p2_k5b_ncu.log
