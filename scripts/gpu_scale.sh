#!/bin/bash
# bench at N GPUs (fused exchange, then NCCL exchange)
set -u
N=${1:-8}
TAG=${2:-s}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 --no-e2e > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err
echo "rc=$?"; cut -c1-260 gpurun_out/bench_${TAG}_n$N.json; tail -3 gpurun_out/bench_${TAG}_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29536 bench.py --gpus $N --steps 20 --warmup 5 --nccl-gather --no-e2e > gpurun_out/bench_${TAG}_n${N}_nccl.json 2> gpurun_out/bench_${TAG}_n${N}_nccl.err
echo "rc=$?"; cut -c1-260 gpurun_out/bench_${TAG}_n${N}_nccl.json
