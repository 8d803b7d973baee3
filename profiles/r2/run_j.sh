#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for v in base new; do
  if [ $v = base ]; then export BLUEICE_B200_LIB=$PWD/blueice_b200/build/variants/lib_base.so; else unset BLUEICE_B200_LIB; fi
  timeout 300 python profiles/r2/prof_k5b_kernel.py 100000000 8 11 64 > gpurun_out/j_k5b_$v.log 2>&1; echo "$v rc=$?"; grep K5B gpurun_out/j_k5b_$v.log
done
