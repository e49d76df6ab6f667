#!/bin/bash
# same-box A/B of two builds: plain kernels (bench lines) and the fused summary epilogue (scripts/dbg_fused_time.py)
set -u
OTHER=${1:-build/ab/libecuda_prev.so}
for LIB in etol_b200/csrc/libecuda.so $OTHER; do
  for J in fd exact; do
   ECUDA_LIB=$PWD/$LIB timeout 300 python bench.py --steps 30 --warmup 5 --jac $J --no-e2e --no-cpu-baseline 2>/dev/null | \
     python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$LIB $J kernel_ms %.4f' % d['roofline']['kernel_ms'])"
  done
  ECUDA_LIB=$PWD/$LIB timeout 200 python scripts/dbg_fused_time.py 2>&1 | tail -1
done
