#!/bin/bash
# round 2, visit B: parity, then A/B of kernel variants (libraries under build/ab/) on one box
set -u
mkdir -p gpurun_out
TAG=${1:-r2b}
echo "== pytest gpu parity"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
for rep in 1 2; do
 for LIB in etol_b200/csrc/libecuda.so $(ls build/ab/*.so 2>/dev/null); do
  ECUDA_LIB=$PWD/$LIB timeout 300 python bench.py --steps 30 --warmup 5 --jac fd --no-e2e --no-cpu-baseline 2>gpurun_out/ab_${TAG}.err | \
   python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$LIB fd kernel_ms %.4f step_ms %.4f frac %.3f' % (d['roofline']['kernel_ms'], d['ms_per_step'], d['roofline']['frac']))"
 done
done
