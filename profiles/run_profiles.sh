#!/bin/bash
# Profiling pass on the GPU box (run through gpurun; each ncu pass only after the plain run exited 0).
#   gpurun --timeout 1500 -- 'bash profiles/run_profiles.sh r1'
# Outputs under gpurun_out/: bench_<tag>.json, launches_<tag>.csv (ncu launch list of the SAME bench command),
# prof_<tag>.ncu-rep (ncu --set full of the K2 kernel in both regimes: config-2 scan and P = 1 over 8 Mi events).
tag=${1:-r1}
set -x
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err || exit 1
python bench.py --steps 2 --warmup 3 --skip-cpu --skip-other > /dev/null 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 2 --warmup 3 --skip-cpu --skip-other > gpurun_out/ncu_launches_${tag}.log 2>&1
python profiles/profile_driver.py 2 > gpurun_out/prof_plain_${tag}.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_unbinned_mma -c 2 -f -o gpurun_out/prof_${tag} \
    python profiles/profile_driver.py 1 > gpurun_out/ncu_full_${tag}.log 2>&1
tail -2 gpurun_out/ncu_full_${tag}.log
# template-space kernels: C5-shaped mixture evaluation (P = 1, 1e8 events) and a C4-shaped toy sweep (1e5 toys)
python profiles/template_profile.py > gpurun_out/tplprof_plain_${tag}.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"k_mixture_partials|k_template_partials" -c 4 -f \
    -o gpurun_out/prof_tpl_${tag} python profiles/template_profile.py > gpurun_out/ncu_full_tpl_${tag}.log 2>&1
tail -2 gpurun_out/ncu_full_tpl_${tag}.log
