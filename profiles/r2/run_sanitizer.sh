#!/bin/bash
# compute-sanitizer memcheck + racecheck over smoke() and a subset of the kernel tests (K2 DMMA, single-launch path,
# tiled K4, K5 / K5b, toys).  Logs -> gpurun_out/ (copied to profiles/ when clean).
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
SEL="single_launch or binned_tiled or (grouped_is_bitwise and 513) or (binned_matches_oracle and 1100) or nan_inf_zero"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_memcheck_smoke.log 2>&1
echo "memcheck smoke rc=$?"; tail -3 gpurun_out/sanitizer_memcheck_smoke.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "$SEL" > gpurun_out/sanitizer_memcheck_kernels.log 2>&1
echo "memcheck kernels rc=$?"; tail -4 gpurun_out/sanitizer_memcheck_kernels.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_template.py tests/test_gpu_toys.py -m gpu -x -q -k "not fit" > gpurun_out/sanitizer_memcheck_template.log 2>&1
echo "memcheck template rc=$?"; tail -4 gpurun_out/sanitizer_memcheck_template.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_racecheck_smoke.log 2>&1
echo "racecheck smoke rc=$?"; tail -3 gpurun_out/sanitizer_racecheck_smoke.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "$SEL" > gpurun_out/sanitizer_racecheck_kernels.log 2>&1
echo "racecheck kernels rc=$?"; tail -4 gpurun_out/sanitizer_racecheck_kernels.log
