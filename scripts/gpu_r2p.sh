#!/bin/bash
# parity + A/B: FD default (one CTA per instance) vs persistent (ECUDA_PERSIST=1); ncu of the persistent kernel
set -u
mkdir -p gpurun_out
TAG=${1:-r2p}
echo "== pytest gpu parity (persistent)"; ECUDA_PERSIST=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
for rep in 1 2; do
 for V in oneshot persist; do
   unset ECUDA_PERSIST
   if [ $V = persist ]; then export ECUDA_PERSIST=1; fi
   timeout 300 python bench.py --steps 30 --warmup 5 --jac fd --no-e2e --no-cpu-baseline --no-extras 2>gpurun_out/ab_${TAG}.err | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$V fd kernel_ms %.4f step_ms %.4f frac %.3f' % (d['roofline']['kernel_ms'], d['ms_per_step'], d['roofline']['frac']))"
 done
done
export ECUDA_PERSIST=1
PROF="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --jac fd"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rows_n -s 4 -c 1 -o gpurun_out/prof_${TAG}_fd -f $PROF > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "ncu rc=$?"
