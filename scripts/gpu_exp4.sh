#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-x}
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "rows:";     KEXP_FLUSH=write+read timeout 300 python scripts/kexp.py 2>&1 | tail -1
echo "columns:";  ECUDA_NO_ROWS=1 KEXP_FLUSH=write+read timeout 300 python scripts/kexp.py 2>&1 | tail -1
for J in fd exact; do
PROF="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --jac $J"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_eval -s 4 -c 1 -o gpurun_out/prof_keval_${TAG}_$J -f $PROF > gpurun_out/ncu_full_${TAG}_$J.log 2>&1
echo "ncu $J rc=$?"
done
