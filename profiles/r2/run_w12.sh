#!/bin/bash
# which DMMA kernel for which batch size: K2 alone at P = 1 .. 1024 for K = 32, 64, 128 (datasets larger than L2)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "wide or k_chunk or mma_kernel_against or batch_shape" > gpurun_out/w12_pytest.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/w12_pytest.log
for cfg in "3 4 1000000" "4 4 500000" "4 8 250000"; do
  BI_MMA_WIDE_MIN_POINTS=100000000 timeout 300 python profiles/r2/wide_probe.py $cfg 1,11,64,256,1024 2>&1 | grep "^{" | tee -a gpurun_out/w12_probe.jsonl
  BI_MMA_WIDE_MIN_TERMS=1 timeout 300 python profiles/r2/wide_probe.py $cfg 1,11,64,256,1024 2>&1 | grep "^{" | tee -a gpurun_out/w12_probe.jsonl
done
