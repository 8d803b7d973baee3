#!/bin/bash
# K4: bulk copies of a tile issued by all warps (one or two lanes each) instead of by the first 40 threads
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "binned" > gpurun_out/k4o_pytest.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/k4o_pytest.log)"
echo "spread   $(timeout 600 python profiles/r2/diag_k4e2e.py 2>&1 | grep DIAG)"
echo "nospread $(BLUEICE_B200_LIB=$GRAFT_REPO_ROOT/blueice_b200/build/variants/lib_k4_nospread.so timeout 600 python profiles/r2/diag_k4e2e.py 2>&1 | grep DIAG)"
echo "spread   $(timeout 600 python profiles/r2/diag_k4e2e.py 2>&1 | grep DIAG)"
