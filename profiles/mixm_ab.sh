set -x
timeout 250 python -m pytest tests/test_gpu_template.py -x -q 2>&1 | tail -4 > gpurun_out/mixm_test.log
for N in 100000000 20000000 2000000; do
  for W in 0 1000000000; do
    TPL_POINTS=7,11,16,64 TPL_WIDE_MIN=$W timeout 90 python profiles/template_bench.py 0 $N 2>&1 | sed "s/^/wide_min=$W /" >> gpurun_out/mixm_ab2.log
  done
  BI_MIX_MMA=0 TPL_POINTS=7,11,16,64 timeout 90 python profiles/template_bench.py 0 $N 2>&1 | sed "s/^/gather /" >> gpurun_out/mixm_ab2.log
done
cat gpurun_out/mixm_test.log gpurun_out/mixm_ab2.log
