#!/bin/bash
# round 2, visit C: parity, A/B new (k_rows_n) vs round-1 kernels for both Jacobian modes, ncu of the exact kernel
set -u
mkdir -p gpurun_out
TAG=${1:-r2c}
echo "== pytest gpu parity"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
for rep in 1 2; do
 for J in exact fd; do
  for V in new old; do
   if [ $V = old ]; then export ECUDA_NO_ROWSN=1; else unset ECUDA_NO_ROWSN; fi
   timeout 300 python bench.py --steps 30 --warmup 5 --jac $J --no-e2e --no-cpu-baseline 2>gpurun_out/ab_${TAG}.err | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$V $J kernel_ms %.4f step_ms %.4f frac %.3f' % (d['roofline']['kernel_ms'], d['ms_per_step'], d['roofline']['frac']))"
  done
 done
done
unset ECUDA_NO_ROWSN
for LIB in $(ls build/ab/*.so 2>/dev/null); do
 for J in exact fd; do
  ECUDA_LIB=$PWD/$LIB timeout 300 python bench.py --steps 30 --warmup 5 --jac $J --no-e2e --no-cpu-baseline 2>>gpurun_out/ab_${TAG}.err | \
   python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$LIB $J kernel_ms %.4f step_ms %.4f frac %.3f' % (d['roofline']['kernel_ms'], d['ms_per_step'], d['roofline']['frac']))"
 done
done
PROF="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --jac exact"
timeout 300 $PROF > gpurun_out/plain_${TAG}.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rows_n -s 4 -c 1 -o gpurun_out/prof_${TAG}_exact -f $PROF > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "ncu rc=$?"
