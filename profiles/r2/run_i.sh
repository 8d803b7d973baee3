#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_template.py -m gpu -x -q > gpurun_out/i_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/i_pytest.log
timeout 300 python profiles/r2/prof_k5b.py 100000000 11 > gpurun_out/i_k5b_p11.log 2>&1; cat gpurun_out/i_k5b_p11.log
timeout 300 python profiles/r2/prof_k5b.py 100000000 64 > gpurun_out/i_k5b_p64.log 2>&1; cat gpurun_out/i_k5b_p64.log
timeout 300 python profiles/r2/prof_k5b.py 12500000 11 > gpurun_out/i_k5b_p11s.log 2>&1; cat gpurun_out/i_k5b_p11s.log
timeout 300 python profiles/r2/prof_k5b.py 100000000 8 > gpurun_out/i_k5b_p8.log 2>&1; cat gpurun_out/i_k5b_p8.log
