#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "--- spread"
CUDA_LAUNCH_BLOCKING=1 timeout 300 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -k "config3" 2>&1 | grep -v "^  File\|Extension modules" | tail -12
echo "--- nospread"
BLUEICE_B200_LIB=$GRAFT_REPO_ROOT/blueice_b200/build/variants/lib_k4_nospread.so CUDA_LAUNCH_BLOCKING=1 timeout 300 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -k "config3" 2>&1 | grep -v "^  File\|Extension modules" | tail -6
