# A/B of the grouped K5b kernels (run through gpurun): tensor-pipe kernel with narrow / wide groups vs the gather kernel
set -x
timeout 250 python -m pytest tests/test_gpu_template.py -x -q 2>&1 | tail -4 > gpurun_out/mixm_test.log
rm -f gpurun_out/mixm_ab2.log
for N in 100000000 20000000 2000000; do
  for W in 0 1000000000; do
    TPL_POINTS=${TPL_POINTS:-7,11,16,64} TPL_WIDE_MIN=$W timeout 90 python profiles/template_bench.py 0 $N 2>&1 | sed "s/^/wide_min=$W /" >> gpurun_out/mixm_ab2.log
  done
  if [ -n "$WITH_GATHER" ]; then
    BI_MIX_MMA=0 TPL_POINTS=${TPL_POINTS:-7,11,16,64} timeout 90 python profiles/template_bench.py 0 $N 2>&1 | sed "s/^/gather /" >> gpurun_out/mixm_ab2.log
  fi
done
cat gpurun_out/mixm_test.log
grep "C5" gpurun_out/mixm_ab2.log | sed 's/C5 mixture bin_major=0: //; s/K=96, //; s/; prepared.*//'
