#!/bin/bash
# compact exact output + pipelined HOST path: GPU tests + bench line with the exact_e2e block, chunk-count sweep
set -u
mkdir -p gpurun_out
TAG=${1:-cx}
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for K in 1 4 8; do
echo "== bench ECUDA_HOST_CHUNKS=$K"; ECUDA_HOST_CHUNKS=$K timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${TAG}_k$K.json 2> gpurun_out/bench_${TAG}_k$K.err
python - <<PY
import json
l=json.loads(open("gpurun_out/bench_${TAG}_k$K.json").read().strip().splitlines()[-1])
print("value %.4e kernel_ms %.4f frac %.3f e2e %.4e frac_of_probe %.3f" % (l["value"], l["roofline"]["kernel_ms"], l["roofline"]["frac"], l["e2e"]["value"], l["e2e"].get("frac_of_d2h_probe", 0)))
x=l.get("exact_e2e",{}); print({k:(v["value"] if isinstance(v,dict) else v) for k,v in x.items() if k!="what"})
PY
tail -3 gpurun_out/bench_${TAG}_k$K.err
done
