#!/bin/bash
# A/B of library variants (etol_b200/csrc/libecuda.so and every build/ab/*.so) on one box, both Jacobian modes,
# kernel-only bench lines, two alternating repetitions. usage: gpu_ab_libs.sh TAG
set -u
mkdir -p gpurun_out
TAG=${1:-ab}
for rep in 1 2; do
 for LIB in etol_b200/csrc/libecuda.so $(ls build/ab/*.so 2>/dev/null); do
  for J in fd exact; do
   ECUDA_LIB=$PWD/$LIB timeout 300 python bench.py --steps 30 --warmup 5 --jac $J --no-e2e --no-cpu-baseline --no-extras 2>>gpurun_out/ab_${TAG}.err | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$LIB $J kernel_ms %.4f step_ms %.4f frac %.3f' % (d['roofline']['kernel_ms'], d['ms_per_step'], d['roofline']['frac']))"
  done
 done
done
tail -3 gpurun_out/ab_${TAG}.err
