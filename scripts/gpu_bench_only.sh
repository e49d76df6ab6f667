#!/bin/bash
# the driver's two bench commands on one GPU
set -u
mkdir -p gpurun_out
TAG=${1:-bo}
nproc > gpurun_out/host_${TAG}.txt; lscpu | grep -E 'Model name|^CPU\(s\)|Thread|Socket' >> gpurun_out/host_${TAG}.txt
echo "== bench (reference arm)"; timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err; tail -c 300 gpurun_out/bench_${TAG}_reference.json; echo
T0=$(date +%s)
echo "== bench (ecuda)"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; cat gpurun_out/bench_${TAG}.json; tail -5 gpurun_out/bench_${TAG}.err
echo "bench wall seconds: $(( $(date +%s) - T0 ))"
