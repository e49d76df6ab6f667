#!/bin/bash
# One GPU-box visit at the end of round 2: smoke, all GPU tests, the driver's two bench commands, the ncu launch list
# of the bench command and one full capture of the dominant kernel (each only after the plain command exited 0).
set -u
mkdir -p gpurun_out
TAG=${1:-r2z}
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/gpu_${TAG}.csv 2>&1
nproc > gpurun_out/host_${TAG}.txt; lscpu | grep -E 'Model name|^CPU\(s\)|Thread|Socket' >> gpurun_out/host_${TAG}.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q -rs 2>&1 | tail -6
echo "== bench (reference arm)"; timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err; tail -c 300 gpurun_out/bench_${TAG}_reference.json; echo
echo "== bench (ecuda)"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${TAG}.json"))
print("value %.4e ms_per_step %.4f kernel_ms %.4f frac %.3f exact %.4f/%.3f e2e %.4e (%.3f of probe) cpu %s" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["exact"]["kernel_ms"], d["exact"]["frac"], d["e2e"]["value"], d["e2e"]["frac_of_d2h_probe"], {k: v for k, v in d["cpu_baseline"].items() if k.endswith("value")}))
print({k: {m: round(v[m]["ms_per_batch"], 4) for m in ("fd", "exact")} for k, v in d["other_configs"].items()})
PY
PROF="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-extras"
timeout 300 $PROF > gpurun_out/plain_${TAG}.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_${TAG}.csv $PROF > gpurun_out/ncu_launch_${TAG}.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rows_n -s 4 -c 1 -o gpurun_out/prof_${TAG}_fd -f $PROF > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "ncu fd rc=$?"
