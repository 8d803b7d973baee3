#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_template.py tests/test_gpu_toys.py -m gpu -x -q > gpurun_out/bm6_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/bm6_pytest.log
for v in "" eb3 eb4 ""; do
  if [ -n "$v" ]; then export BLUEICE_B200_LIB=$GRAFT_REPO_ROOT/blueice_b200/build/variants/lib_$v.so; else unset BLUEICE_B200_LIB; fi
  echo "variant: ${v:-default} $(timeout 600 python profiles/r2/diag1.py k5 2>&1 | grep DIAG)"
done
