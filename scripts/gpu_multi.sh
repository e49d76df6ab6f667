#!/bin/bash
# N-GPU bench through torchrun, the way the driver launches it
set -u
N=${1:-2}
TAG=${2:-m}
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err
echo "rc=$?"; cat gpurun_out/bench_${TAG}_n$N.json; tail -5 gpurun_out/bench_${TAG}_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 10 --warmup 3 --gather full --no-e2e > gpurun_out/bench_${TAG}_n${N}_full.json 2> gpurun_out/bench_${TAG}_n${N}_full.err
echo "rc=$?"; cat gpurun_out/bench_${TAG}_n${N}_full.json; tail -5 gpurun_out/bench_${TAG}_n${N}_full.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 bench.py --impl reference --gpus $N --steps 2 --warmup 1 | tail -c 400
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29536 bench.py --gpus $N --steps 20 --warmup 5 --nccl-gather --no-e2e > gpurun_out/bench_${TAG}_n${N}_nccl.json 2> gpurun_out/bench_${TAG}_n${N}_nccl.err; echo "rc=$?"; cat gpurun_out/bench_${TAG}_n${N}_nccl.json | cut -c1-200; tail -3 gpurun_out/bench_${TAG}_n${N}_nccl.err
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
