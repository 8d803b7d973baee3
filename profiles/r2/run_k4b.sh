#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_api.py -m gpu -x -q -k "binned or Binned or beeston or bb" > gpurun_out/k4b_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/k4b_pytest.log
for tile in 128 256; do
  BI_BINNED_TILE=$tile timeout 600 python profiles/r2/diag1.py k4 > gpurun_out/k4b_diag_$tile.log 2>&1; echo "tile $tile: $(grep DIAG gpurun_out/k4b_diag_$tile.log || tail -5 gpurun_out/k4b_diag_$tile.log)"
done
timeout 600 python profiles/r2/diag1.py k4 > gpurun_out/k4b_diag_auto.log 2>&1; echo "auto: $(grep DIAG gpurun_out/k4b_diag_auto.log)"
# launch list (kernel durations) of the P = 1 and P = 256 evaluations: 6 evaluations each
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_binned|k_canonical" -c 40 --csv --log-file gpurun_out/k4b_p1_launches.csv \
    python profiles/r2/diag1.py k4p1 > /dev/null 2>&1
tail -12 gpurun_out/k4b_p1_launches.csv | cut -d, -f5,14- | cut -c1-150
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_binned|k_canonical" -c 40 --csv --log-file gpurun_out/k4b_p256_launches.csv \
    python profiles/r2/diag1.py k4big > /dev/null 2>&1
tail -12 gpurun_out/k4b_p256_launches.csv | cut -d, -f5,14- | cut -c1-150
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_binned_tile -c 2 -f -o gpurun_out/r2_k4_tiled_p256 \
    python profiles/r2/diag1.py k4big > gpurun_out/ncu_k4_tiled_p256.log 2>&1
tail -2 gpurun_out/ncu_k4_tiled_p256.log
