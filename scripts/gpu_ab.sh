#!/bin/bash
# A/B of two builds of libecuda.so on the bench workload (kernel-only lines)
set -u
mkdir -p gpurun_out
TAG=${1:-ab}
echo "== pytest -m gpu (default lib)"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for LIB in libecuda.so libecuda_3cta.so; do
 for J in fd exact; do
  ECUDA_LIB=$PWD/etol_b200/csrc/$LIB timeout 300 python bench.py --steps 30 --warmup 5 --jac $J --no-e2e --no-cpu-baseline > gpurun_out/bench_${TAG}_${LIB}_$J.json 2> gpurun_out/bench_${TAG}_${LIB}_$J.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${TAG}_${LIB}_$J.json"))
    print("$LIB $J value %.4e evals/s  kernel_ms %.4f  frac %.3f" % (d["value"], d["roofline"]["kernel_ms"], d["roofline"]["frac"]))
except Exception as e:
    print("bench $LIB $J failed", e); print(open("gpurun_out/bench_${TAG}_${LIB}_$J.err").read()[-1500:])
PY
 done
done
echo "== parity with 3cta lib"; ECUDA_LIB=$PWD/etol_b200/csrc/libecuda_3cta.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "values_match or full_size" 2>&1 | tail -3
PROF="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 300 $PROF > gpurun_out/plain_${TAG}.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_eval -s 4 -c 1 -o gpurun_out/prof_keval_${TAG} -f $PROF > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "ncu rc=$?"
