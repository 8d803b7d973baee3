#!/bin/bash
# chunk length of the K-chunk kernel: 32 terms x 2 stages (default) against 16 terms x 2 / 3 / 4 stages
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for lib in "" blueice_b200/build/variants/lib_wide_k16s2.so blueice_b200/build/variants/lib_wide_k16s3.so blueice_b200/build/variants/lib_wide_k16s4.so; do
  echo "lib=$lib" | tee -a gpurun_out/w15_probe.jsonl
  if [ -n "$lib" ]; then
    BLUEICE_B200_LIB=$lib timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "wide_contraction or mma_kernel_against" 2>&1 | tail -1
  fi
  for cfg in "5 5 50000 4096" "4 4 100000 4096" "4 12 50000 4096"; do
    BLUEICE_B200_LIB=$lib BI_WIDE_VERBOSE=1 timeout 300 python profiles/r2/wide_probe.py $cfg 2>&1 | grep "^{\|CTAs per" | tee -a gpurun_out/w15_probe.jsonl
  done
done
