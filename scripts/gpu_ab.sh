#!/bin/bash
# A/B of two builds of libecuda.so on the bench workload, same box (kernel-only lines, alternating)
# usage: gpu_ab.sh <other .so, path relative to the repo>   (the default library is etol_b200/csrc/libecuda.so)
set -u
OTHER=${1:-build/ab/libecuda_prev.so}
for rep in 1 2; do
 for LIB in etol_b200/csrc/libecuda.so $OTHER; do
  for J in fd exact; do
   ECUDA_LIB=$PWD/$LIB timeout 300 python bench.py --steps 30 --warmup 5 --jac $J --no-e2e --no-cpu-baseline 2>/dev/null | \
     python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$LIB $J kernel_ms %.4f step_ms %.4f' % (d['roofline']['kernel_ms'], d['ms_per_step']))"
  done
 done
done
