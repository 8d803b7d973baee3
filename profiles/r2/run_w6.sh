#!/bin/bash
# full ncu capture of the K-chunk kernel v2 on a 4096-point scan with 160 contraction terms
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python profiles/r2/wide_probe.py 5 5 50000 4096 > gpurun_out/w6_plain.log 2>&1 || exit 1
tail -1 gpurun_out/w6_plain.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_unbinned_mma_wide -s 3 -c 1 -o gpurun_out/w6_wide -f python profiles/r2/wide_probe.py 5 5 50000 4096 > gpurun_out/w6_ncu2.log 2>&1
echo "ncu rc=$?"
