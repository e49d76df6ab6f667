#!/bin/bash
# round 2, visit H: GPU tests (all), bench A/B new vs round-1 kernels, ncu of the streaming exact kernel
set -u
mkdir -p gpurun_out
TAG=${1:-r2h}
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for rep in 1 2; do
 for J in fd exact; do
  for V in new old rowsexact; do
   unset ECUDA_NO_ROWSN ECUDA_ROWS_EXACT
   if [ $V = old ]; then export ECUDA_NO_ROWSN=1; fi
   if [ $V = rowsexact ]; then if [ $J = fd ]; then continue; fi; export ECUDA_ROWS_EXACT=1; fi
   timeout 300 python bench.py --steps 30 --warmup 5 --jac $J --no-e2e --no-cpu-baseline --no-extras 2>gpurun_out/ab_${TAG}.err | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$V $J kernel_ms %.4f step_ms %.4f frac %.3f' % (d['roofline']['kernel_ms'], d['ms_per_step'], d['roofline']['frac']))"
  done
 done
done
unset ECUDA_NO_ROWSN ECUDA_ROWS_EXACT
PROF="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --jac exact"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stream -s 4 -c 1 -o gpurun_out/prof_${TAG}_exact -f $PROF > gpurun_out/ncu_full_${TAG}_exact.log 2>&1
echo "ncu exact rc=$?"
