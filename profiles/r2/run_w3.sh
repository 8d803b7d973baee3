#!/bin/bash
# launch list and full ncu capture of the K-chunk kernel on a 4096-point scan with 160 contraction terms
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python profiles/r2/wide_probe.py 5 5 50000 4096 > gpurun_out/w3_plain.log 2>&1 || exit 1
tail -1 gpurun_out/w3_plain.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/w3_launches.csv python profiles/r2/wide_probe.py 5 5 50000 4096 > gpurun_out/w3_ncu1.log 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open('gpurun_out/w3_launches.csv')) if len(r) > 10]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
for r in rows[-12:]: print(r[ki][:60], r[vi])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_unbinned_mma_wide -s 3 -c 1 -o gpurun_out/w3_wide -f python profiles/r2/wide_probe.py 5 5 50000 4096 > gpurun_out/w3_ncu2.log 2>&1
echo "ncu rc=$?"
