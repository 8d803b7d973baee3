#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for v in "" c1024_4 c1024_3 c4096_t512; do
  if [ -n "$v" ]; then export BLUEICE_B200_LIB=$GRAFT_REPO_ROOT/blueice_b200/build/variants/lib_$v.so; fi
  echo "variant: ${v:-default} $(timeout 600 python profiles/r2/diag1.py k5 2>&1 | grep DIAG)"
done
