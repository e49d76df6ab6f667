#!/bin/bash
# one ncu --set full capture of the k_rows_n kernel of one Jacobian mode. usage: gpu_ncu1.sh TAG fd|exact
set -u
mkdir -p gpurun_out
TAG=${1:-n}; J=${2:-fd}
PROF="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --jac $J"
timeout 300 $PROF > gpurun_out/plain_${TAG}.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rows_n -s 4 -c 1 -o gpurun_out/prof_${TAG}_$J -f $PROF > gpurun_out/ncu_full_${TAG}_$J.log 2>&1
echo "ncu $J rc=$?"; ls -la gpurun_out
