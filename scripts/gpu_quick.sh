#!/bin/bash
# Short GPU visit while iterating on a kernel: parity tests, kernel-only bench lines, one ncu capture.
# usage: gpu_quick.sh TAG [jac-mode-for-ncu]
set -u
mkdir -p gpurun_out
TAG=${1:-q}
NJ=${2:-fd}
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for J in fd exact; do
  timeout 300 python bench.py --steps 20 --warmup 5 --jac $J --no-e2e --no-cpu-baseline > gpurun_out/bench_${TAG}_$J.json 2> gpurun_out/bench_${TAG}_$J.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${TAG}_$J.json"))
    print("$J value %.4e evals/s  kernel_ms %.4f  frac %.3f  clocks %s" % (d["value"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["clocks"]))
except Exception as e:
    print("bench $J failed", e); print(open("gpurun_out/bench_${TAG}_$J.err").read()[-1500:])
PY
done
PROF="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --jac $NJ"
timeout 300 $PROF > gpurun_out/plain_${TAG}.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_eval -s 4 -c 1 -o gpurun_out/prof_keval_${TAG} -f $PROF > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "ncu rc=$?"
