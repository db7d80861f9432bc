# 8-GPU (or N-GPU) re-measurement after the panel-encoding / rotation changes: bench under torchrun + multi_check
set -x
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n${N}_r02.json 2> gpurun_out/bench_n${N}_r02.err; echo rc=$?
tail -c 200 gpurun_out/bench_n${N}_r02.err
python tools/multi_check.py --gpus $N --steps 4 > gpurun_out/multi_check_n${N}.json 2> gpurun_out/multi_check_n${N}.err; echo rc=$?
tail -c 200 gpurun_out/multi_check_n${N}.err
