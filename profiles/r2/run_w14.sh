#!/bin/bash
# deep ring (6 stages, 1 CTA / SM) against the 2-stage form at a handful of points
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "wide or k_chunk or mma_kernel_against or sourcewise" > gpurun_out/w14_pytest.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/w14_pytest.log
for cfg in "5 5 50000" "5 5 200000" "5 5 800000" "4 4 500000"; do
  for deep in 0 1; do
    echo "deep=$deep cfg=$cfg" | tee -a gpurun_out/w14_probe.jsonl
    BI_WIDE_DEEP=$deep BI_WIDE_VERBOSE=1 timeout 300 python profiles/r2/wide_probe.py $cfg 1,2,4,11 2>&1 | grep "^{" | tee -a gpurun_out/w14_probe.jsonl
  done
done
