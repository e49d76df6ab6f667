#!/bin/bash
# one-GPU round check: smoke, all GPU tests, both bench arms (the driver's commands), launch list
set -u
mkdir -p gpurun_out
TAG=${1:-b1}
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/gpu_${TAG}.csv 2>&1
nproc > gpurun_out/host_${TAG}.txt; lscpu | grep -E 'Model name|^CPU\(s\)|Thread|Socket' >> gpurun_out/host_${TAG}.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== bench (reference arm)"; timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err; tail -c 400 gpurun_out/bench_${TAG}_reference.json; echo
echo "== bench (ecuda)"; /usr/bin/time -v timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; cat gpurun_out/bench_${TAG}.json; grep -E "Elapsed|Error|Traceback" -A3 gpurun_out/bench_${TAG}.err | tail -12
PROF="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-extras"
timeout 300 $PROF > gpurun_out/plain_${TAG}.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_${TAG}.csv $PROF > gpurun_out/ncu_launch_${TAG}.log 2>&1
echo "launch list rc=$?"
