#!/bin/bash
# crossover of the per-warp-ring kernel and the K-chunk kernel over the contraction length
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "k_chunk" > gpurun_out/w2_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/w2_pytest.log
for cfg in "2 3 100000 4096" "2 4 100000 4096" "3 3 100000 4096" "3 4 100000 4096" "4 3 100000 4096" "4 4 100000 4096" "3 10 100000 4096"; do
  timeout 300 python profiles/r2/wide_probe.py $cfg 2>&1 | tail -1 | tee -a gpurun_out/w2_probe.jsonl
  BI_MMA_WIDE_MIN_TERMS=1 timeout 300 python profiles/r2/wide_probe.py $cfg 2>&1 | tail -1 | tee -a gpurun_out/w2_probe.jsonl
done
