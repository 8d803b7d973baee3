#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python profiles/r2/prof_c5_host.py > gpurun_out/h_c5host.log 2>&1; echo "c5host rc=$?"; grep C5HOST gpurun_out/h_c5host.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/h_c5_launches.csv python profiles/r2/prof_k5b.py 12500000 11 > gpurun_out/h_c5_ncu.log 2>&1; echo "ncu rc=$?"
for v in k2u1 k2u2; do
BLUEICE_B200_LIB=$PWD/blueice_b200/build/variants/lib_$v.so timeout 600 python bench.py --steps 30 --warmup 5 --skip-cpu --skip-other > gpurun_out/h_bench_$v.json 2> gpurun_out/h_bench_$v.err; echo "$v bench rc=$?"
python - $v <<'PY'
import json, sys
d = json.loads(open('gpurun_out/h_bench_%s.json' % sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "value %.4e ms %.4f e2e %.4e k2_ms %.4f frac %.4f stream %.4f" % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['ms'], d['roofline']['frac'], d['roofline_stream']['frac']))
PY
done
