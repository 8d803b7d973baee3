#!/bin/bash
# ring depth of the K-chunk kernel: 2 stages (3 CTAs / SM), 3 stages (2), 4 stages (1) over batch sizes
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for lib in "" blueice_b200/build/variants/lib_wide_s3.so blueice_b200/build/variants/lib_wide_s4.so; do
  for cfg in "4 4 500000" "5 5 200000"; do
    echo "lib=$lib cfg=$cfg" | tee -a gpurun_out/w13_probe.jsonl
    BLUEICE_B200_LIB=$lib BI_WIDE_VERBOSE=1 timeout 300 python profiles/r2/wide_probe.py $cfg 1,11,64,1024,4096 2>&1 | grep "^{\|CTAs per SM" | tee -a gpurun_out/w13_probe.jsonl
  done
done
