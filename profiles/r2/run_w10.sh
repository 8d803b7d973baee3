#!/bin/bash
# crossover of the two DMMA kernels with the final K-chunk kernel
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for cfg in "2 4 100000 4096" "3 3 100000 4096" "3 4 100000 4096" "3 5 100000 4096" "4 3 100000 4096"; do
  timeout 300 python profiles/r2/wide_probe.py $cfg 2>&1 | tail -1 | tee -a gpurun_out/w10_probe.jsonl
  BI_MMA_WIDE_MIN_TERMS=1 timeout 300 python profiles/r2/wide_probe.py $cfg 2>&1 | tail -1 | tee -a gpurun_out/w10_probe.jsonl
done
