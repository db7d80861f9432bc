# Round-2 run on 8 GPUs of one box: multi-GPU context tests, the bench under torchrun (value: one rank per GPU; e2e: one
# C-ABI call driving all 8 GPUs from rank 0; configs[3] and configs[4] at N = 8), and the one-process multi-GPU timings.
set -x
N=${1:-8}
python -m pytest tests/test_gpu_multi.py tests/test_cabi_smoke.py -m gpu -x -q 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n${N}_r02.json 2> gpurun_out/bench_n${N}_r02.err; echo rc=$?
tail -c 300 gpurun_out/bench_n${N}_r02.err
python tools/multi_check.py --gpus $N --steps 4 > gpurun_out/multi_check_n${N}.json 2> gpurun_out/multi_check_n${N}.err; echo rc=$?
tail -c 300 gpurun_out/multi_check_n${N}.err
