#!/bin/bash
# A/B of environment switches on one box: bench kernel time, FD and exact, interleaved repeats
# usage: gpu_ab_env.sh TAG "ENV1=a" "ENV2=b" ...   (an empty string "" = defaults)
set -u
mkdir -p gpurun_out
TAG=$1; shift
for rep in 1 2; do
for E in "$@"; do
for J in fd exact; do
  env $E timeout 300 python bench.py --jac $J --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/ab_${TAG}.json 2> gpurun_out/ab_${TAG}.err
  python - "$E" $J <<PY
import json,sys
try:
    l=json.loads(open("gpurun_out/ab_${TAG}.json").read().strip().splitlines()[-1])
    print("%-28s %-5s kernel_ms %.4f step_ms %.4f frac %.3f launches %d" % (sys.argv[1] or "default", sys.argv[2], l["roofline"]["kernel_ms"], l["ms_per_step"], l["roofline"]["frac"], l["gpu_launches"]))
except Exception as e:
    print(sys.argv[1], sys.argv[2], "FAILED", e); print(open("gpurun_out/ab_${TAG}.err").read()[-600:])
PY
done; done; done
