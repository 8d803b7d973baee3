#!/bin/bash
# K-chunk kernel after the register-cap change: tests of the kernel, timings
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "wide or k_chunk or mma_kernel_against" > gpurun_out/w4_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/w4_pytest.log
for cfg in "5 5 50000 4096" "4 12 50000 4096" "3 40 50000 2048" "4 4 100000 4096" "3 10 100000 4096"; do
  timeout 300 python profiles/r2/wide_probe.py $cfg 2>&1 | tail -1 | tee -a gpurun_out/w4_probe.jsonl
done
