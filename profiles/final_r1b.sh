# final pass of the round (through gpurun): GPU tests, smoke, bench line, ncu capture of the tensor-pipe K5b kernel
set -x
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/pytest_r1b.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r1b.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1b.json 2> gpurun_out/bench_r1b.err
TPL_POINTS=7,11 timeout 100 python profiles/template_bench.py 0 100000000 > gpurun_out/mixm_plain_r1b.log 2>&1 && \
TPL_POINTS=7,11 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_mixture_partials_mma \
    --launch-skip 3 --launch-count 2 -f -o gpurun_out/prof_mixm_r1b python profiles/template_bench.py 0 100000000 > gpurun_out/ncu_mixm_r1b.log 2>&1
tail -2 gpurun_out/ncu_mixm_r1b.log; cat gpurun_out/pytest_r1b.log gpurun_out/mixm_plain_r1b.log; tail -3 gpurun_out/smoke_r1b.log; tail -c 300 gpurun_out/bench_r1b.err
