#!/bin/bash
# One GPU-box visit: smoke, GPU parity tests, bench (both arms), then ncu launch list + full capture
# of the dominant kernel. Everything it produces lands in gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-r1}
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/gpu_${TAG}.csv 2>&1
nproc > gpurun_out/host_${TAG}.txt; lscpu | grep -E 'Model name|^CPU\(s\)|Thread|Socket' >> gpurun_out/host_${TAG}.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== bench (reference arm)"; timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err; tail -c 600 gpurun_out/bench_${TAG}_reference.json
echo "== bench (ecuda)"; timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; cat gpurun_out/bench_${TAG}.json; tail -5 gpurun_out/bench_${TAG}.err
echo "== bench exact-mode"; timeout 300 python bench.py --steps 20 --warmup 5 --jac exact --no-e2e --no-cpu-baseline > gpurun_out/bench_${TAG}_exact.json 2>> gpurun_out/bench_${TAG}.err; cat gpurun_out/bench_${TAG}_exact.json
PROF="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
echo "== ncu"
timeout 300 $PROF > gpurun_out/plain_${TAG}.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_${TAG}.csv $PROF > gpurun_out/ncu_launch_${TAG}.log 2>&1
echo "launch list rc=$?"
timeout 300 $PROF > gpurun_out/plain2_${TAG}.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_eval -s 4 -c 2 -o gpurun_out/prof_keval_${TAG} -f $PROF > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "full capture rc=$?"
echo "== other configurations / user models / C3 slices"
timeout 600 python scripts/config_sweep.py > gpurun_out/config_sweep_${TAG}.json 2> gpurun_out/config_sweep_${TAG}.err; tail -3 gpurun_out/config_sweep_${TAG}.err
timeout 600 python scripts/user_model_time.py > gpurun_out/user_model_time_${TAG}.log 2>&1; tail -3 gpurun_out/user_model_time_${TAG}.log
timeout 300 python scripts/run_examples.py > gpurun_out/examples_run_${TAG}.txt 2>&1; grep -E "rc |Score|mesh:" gpurun_out/examples_run_${TAG}.txt
ls -la gpurun_out | tail -20
