# Round-2 run B (2 GPUs): tests, the L2 residency experiment, the permutation step's launch list, bench under torchrun
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/tests_r02b.log; tail -3 gpurun_out/tests_r02b.log
for mode in 0 1 2; do
  BLMM_B200_L2_PERSIST=$mode BLMM_B200_TRACE=1 python tools/quick_bxd.py > gpurun_out/l2_mode${mode}.log 2>&1
  BLMM_B200_L2_PERSIST=$mode ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:scan_kernel -s 2 -c 1 --csv --log-file gpurun_out/l2_mode${mode}_dram.csv python tools/quick_bxd.py > /dev/null 2>&1
done
grep -h "alt-grid:" gpurun_out/l2_mode*.log | awk 'NR%4==0'
python tools/quick_perms.py > gpurun_out/quick_perms.log 2>&1; tail -2 gpurun_out/quick_perms.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_perms_r02.csv python tools/quick_perms.py 4 > /dev/null 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2_r02.json 2> gpurun_out/bench_n2_r02.err; echo rc=$?
tail -c 400 gpurun_out/bench_n2_r02.err
