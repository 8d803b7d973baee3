#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "--- spread (default lib): api + fullsize"
timeout 600 python -m pytest tests/test_gpu_api.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/k4q_a.log 2>&1; echo "rc=$?"; grep -v "^  File\|Extension modules" gpurun_out/k4q_a.log | tail -5 | cut -c1-200
echo "--- nospread lib: api + fullsize"
BLUEICE_B200_LIB=$GRAFT_REPO_ROOT/blueice_b200/build/variants/lib_k4_nospread.so timeout 600 python -m pytest tests/test_gpu_api.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/k4q_b.log 2>&1; echo "rc=$?"; grep -v "^  File\|Extension modules" gpurun_out/k4q_b.log | tail -5 | cut -c1-200
