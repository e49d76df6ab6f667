#!/bin/bash
# N GPUs of one box: multi-GPU parity test, then the driver's bench commands at that N (C5 sharded)
set -u
mkdir -p gpurun_out
N=${1:-2}; TAG=${2:-m}
nvidia-smi topo -m > gpurun_out/topo_${TAG}.txt 2>&1
echo "== multi-GPU parity"; timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -rs 2>&1 | tail -5
echo "== bench --gpus $N"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${TAG}_n$N.json"))
    print("value %.4e ms_per_step %.4f scaling %s" % (d["value"], d["ms_per_step"], d["scaling"]))
    print("workload:", d["config"]["workload"])
    print("e2e:", d.get("e2e")); print("gather_parity:", d.get("gather_parity")); print("gather_full:", d.get("gather_full")); print("d2h_probe:", d.get("d2h_probe")); print("exact:", d.get("exact")); print("roofline frac", d["roofline"]["frac"], d["roofline"]["kernel_ms"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_${TAG}_n$N.err").read()[-3000:])
PY
echo "== reference arm at N=$N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>/dev/null | tail -c 600

