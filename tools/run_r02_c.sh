# Round-2 run C (1 GPU): tests with the new defaults, two-tiles-per-turn experiment on the one-k scans, Brent staging
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/tests_r02c.log; tail -2 gpurun_out/tests_r02c.log
for tpt in 1 2; do
  BLMM_B200_SCAN_TPT=$tpt python tools/quick_bxd.py > gpurun_out/tpt${tpt}_bxd.log 2>&1; grep "null-grid" gpurun_out/tpt${tpt}_bxd.log | tail -2
  BLMM_B200_SCAN_TPT=$tpt python tools/quick_perms.py > gpurun_out/tpt${tpt}_perms.log 2>&1; tail -2 gpurun_out/tpt${tpt}_perms.log
done
BLMM_B200_SCAN_TPT=2 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge.py tests/test_gpu_random_shapes.py tests/test_gpu_parity_full.py -m gpu -x -q 2>&1 | tail -3
python tools/quick_exact.py 79 7321 35554 0 4 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_perms_r02c.csv python tools/quick_perms.py 3 > /dev/null 2>&1
grep -E "fit_h2|scan_kernel" gpurun_out/launches_perms_r02c.csv | tail -4 | awk -F'","' '{print $5, $NF}' | cut -c1-120
