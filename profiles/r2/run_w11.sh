#!/bin/bash
# few-point regime (minimiser steps): the two DMMA kernels at P = 1 and P = 11 over datasets larger than L2
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for cfg in "3 4 1000000 1" "3 4 1000000 11" "4 4 500000 1" "4 4 500000 11" "3 3 1000000 1"; do
  timeout 300 python profiles/r2/wide_probe.py $cfg 2>&1 | tail -1 | tee -a gpurun_out/w11_probe.jsonl
  BI_MMA_WIDE_MIN_TERMS=1 timeout 300 python profiles/r2/wide_probe.py $cfg 2>&1 | tail -1 | tee -a gpurun_out/w11_probe.jsonl
done
for cfg in "5 5 200000 1" "5 5 200000 11"; do
  timeout 300 python profiles/r2/wide_probe.py $cfg stream 2>&1 | tail -1 | tee -a gpurun_out/w11_probe.jsonl
done
