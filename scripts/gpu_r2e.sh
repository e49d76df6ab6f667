#!/bin/bash
# round 2, visit E: GPU parity of the staged-store kernels, bench both modes (new vs round-1 kernels), ncu of both
set -u
mkdir -p gpurun_out
TAG=${1:-r2e}
echo "== pytest gpu parity"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -6
for rep in 1 2; do
 for J in fd exact; do
  for V in new old; do
   if [ $V = old ]; then export ECUDA_NO_ROWSN=1; else unset ECUDA_NO_ROWSN; fi
   timeout 300 python bench.py --steps 30 --warmup 5 --jac $J --no-e2e --no-cpu-baseline --no-extras 2>gpurun_out/ab_${TAG}.err | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$V $J kernel_ms %.4f step_ms %.4f frac %.3f' % (d['roofline']['kernel_ms'], d['ms_per_step'], d['roofline']['frac']))"
  done
 done
done
unset ECUDA_NO_ROWSN
for J in fd exact; do
PROF="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --jac $J"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rows_n -s 4 -c 1 -o gpurun_out/prof_${TAG}_$J -f $PROF > gpurun_out/ncu_full_${TAG}_$J.log 2>&1
echo "ncu $J rc=$?"
done
