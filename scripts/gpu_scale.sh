#!/bin/bash
# 8-GPU box: the driver's scaling commands at N = 4 and 8 (C5 sharded), multi-GPU parity test, topology
set -u
mkdir -p gpurun_out
TAG=${1:-sc}
nvidia-smi topo -m > gpurun_out/topo_${TAG}.txt 2>&1
echo "== multi-GPU parity"; timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -rs 2>&1 | tail -4
for N in 8 4; do
echo "== bench --gpus $N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${TAG}_n$N.json"))
    print("N=$N value %.4e ms_per_step %.4f scaling %s frac %.3f kernel_ms %.4f" % (d["value"], d["ms_per_step"], d["scaling"], d["roofline"]["frac"], d["roofline"]["kernel_ms"]))
    print(" e2e:", {k: d["e2e"][k] for k in ("value", "frac_of_d2h_probe")} if d.get("e2e") else None)
    print(" gather_parity:", d.get("gather_parity"))
    gf = d.get("gather_full") or {}
    print(" gather_full:", {k: gf.get(k) for k in ("ms", "GBps_per_rank", "frac_of_nvlink_900GBps", "parity", "error")})
    print(" d2h_probe:", d.get("d2h_probe", {}).get("GBps_per_gpu"), " exact:", (d.get("exact") or {}).get("frac"))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_${TAG}_n$N.err").read()[-2500:])
PY
done
