#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_template.py -m gpu -x -q > gpurun_out/j_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/j_pytest.log
for v in ${VARIANTS:-new mc4}; do
  if [ $v = new ]; then unset BLUEICE_B200_LIB; else export BLUEICE_B200_LIB=$PWD/blueice_b200/build/variants/lib_$v.so; fi
  timeout 300 python profiles/r2/prof_k5b_kernel.py 100000000 11 64 > gpurun_out/j_k5b_$v.log 2>&1; echo "$v rc=$?"; grep K5B gpurun_out/j_k5b_$v.log
done
