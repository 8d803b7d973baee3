#!/bin/bash
# first run of the K-chunk kernel: its tests, then timings
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "wide or k_chunk or mma_kernel_against or stream_matches" > gpurun_out/w1_pytest.log 2>&1
echo "pytest rc=$?"; tail -25 gpurun_out/w1_pytest.log
for cfg in "5 5 50000 4096 stream" "4 12 50000 4096" "5 8 20000 4096" "3 40 50000 2048"; do
  timeout 300 python profiles/r2/wide_probe.py $cfg 2>&1 | tail -1 | tee -a gpurun_out/w1_probe.jsonl
done
for cfg in "3 12 100000 4096" "4 8 50000 4096" "2 2 99957 4096"; do
  timeout 300 python profiles/r2/wide_probe.py $cfg 2>&1 | tail -1 | tee -a gpurun_out/w1_probe.jsonl
  BI_MMA_WIDE_MIN_TERMS=1 timeout 300 python profiles/r2/wide_probe.py $cfg 2>&1 | tail -1 | tee -a gpurun_out/w1_probe.jsonl
done
