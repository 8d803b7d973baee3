#!/bin/bash
# K-chunk kernel v4 (coefficients packed chunk-major, one producer warp, 2 stages, 3 CTAs / SM): tests, timings, 3-stage variant
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "wide or k_chunk or mma_kernel_against" > gpurun_out/w9_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/w9_pytest.log
for cfg in "5 5 50000 4096" "4 12 50000 4096" "3 40 50000 2048" "4 4 100000 4096" "3 10 100000 4096" "4 3 100000 4096"; do
  BI_WIDE_VERBOSE=1 timeout 300 python profiles/r2/wide_probe.py $cfg 2>&1 | tail -2 | tee -a gpurun_out/w9_probe.jsonl
done
echo "--- 3 stages"
for cfg in "5 5 50000 4096" "4 4 100000 4096"; do
  BI_WIDE_VERBOSE=1 BLUEICE_B200_LIB=blueice_b200/build/variants/lib_wide_s3.so timeout 300 python profiles/r2/wide_probe.py $cfg 2>&1 | tail -2 | tee -a gpurun_out/w9_probe.jsonl
done
