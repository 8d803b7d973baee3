#!/bin/bash
# K-chunk kernel: + only the live coefficient rows copied
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_api.py -m gpu -x -q -k "wide or k_chunk or mma_kernel_against or sourcewise or long_contraction" > gpurun_out/w17_pytest.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/w17_pytest.log
for cfg in "5 5 200000" "4 4 500000"; do
  timeout 300 python profiles/r2/wide_probe.py $cfg 1,11,64,256,1024,4096 2>&1 | grep "^{" | tee -a gpurun_out/w17_probe.jsonl
done
