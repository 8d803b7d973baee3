#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for f in 0 1; do
BI_BINNED_FUSED=$f timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/k4g_p1_f$f.csv \
    python profiles/r2/diag1.py k4p1 > gpurun_out/k4g_l$f.log 2>&1
echo "fused=$f"; tail -8 gpurun_out/k4g_p1_f$f.csv | cut -d, -f5,8,9,15 | cut -c1-150
done
