#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for i in 1 2; do
  TORCH_SHOW_CPP_STACKTRACES=1 timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/k4r_$i.log 2>&1; echo "run $i rc=$?"
  grep -v "^  File\|Extension modules" gpurun_out/k4r_$i.log | tail -4 | cut -c1-300
  grep -n "CUDA error\|terminate\|what()\|illegal\|unspecified" gpurun_out/k4r_$i.log | head -5
done
