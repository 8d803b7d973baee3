#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for f in 4 5; do
  export BI_BINNED_FEW=$f
  timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "binned" > gpurun_out/k4n_pytest$f.log 2>&1
  echo "few=$f pytest rc=$? $(tail -1 gpurun_out/k4n_pytest$f.log)"
  echo "few=$f $(timeout 600 python profiles/r2/diag_k4e2e.py 2>&1 | grep DIAG)"
done
