#!/bin/bash
# round 2, visit A: parity of the N-specialised FD kernel, A/B against the round-1 kernel, one ncu capture
set -u
mkdir -p gpurun_out
TAG=${1:-r2a}
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/gpu_${TAG}.csv 2>&1
echo "== pytest gpu parity"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
for rep in 1 2; do
 for V in new old; do
  if [ $V = old ]; then export ECUDA_NO_ROWSN=1; else unset ECUDA_NO_ROWSN; fi
  timeout 300 python bench.py --steps 30 --warmup 5 --jac fd --no-e2e --no-cpu-baseline 2>gpurun_out/ab_${TAG}.err | \
   python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$V fd kernel_ms %.4f step_ms %.4f frac %.3f' % (d['roofline']['kernel_ms'], d['ms_per_step'], d['roofline']['frac']))"
 done
done
unset ECUDA_NO_ROWSN
PROF="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --jac fd"
timeout 300 $PROF > gpurun_out/plain_${TAG}.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rows_n -s 4 -c 1 -o gpurun_out/prof_${TAG}_fd -f $PROF > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "ncu rc=$?"
