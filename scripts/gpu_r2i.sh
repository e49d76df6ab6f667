#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r2i}
echo "== pytest gpu parity"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
for rep in 1 2; do
 for V in new old; do
   unset ECUDA_NO_ROWSN
   if [ $V = old ]; then export ECUDA_NO_ROWSN=1; fi
   timeout 300 python bench.py --steps 30 --warmup 5 --jac exact --no-e2e --no-cpu-baseline --no-extras 2>gpurun_out/ab_${TAG}.err | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$V exact kernel_ms %.4f step_ms %.4f frac %.3f' % (d['roofline']['kernel_ms'], d['ms_per_step'], d['roofline']['frac']))"
 done
done
unset ECUDA_NO_ROWSN
PROF="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --jac exact"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stream -s 4 -c 1 -o gpurun_out/prof_${TAG}_exact -f $PROF > gpurun_out/ncu_full_${TAG}_exact.log 2>&1
echo "ncu exact rc=$?"
