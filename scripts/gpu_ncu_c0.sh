#!/bin/bash
# ncu --set full of one launch of the FD kernel of C0 (reference VGP shape, B=4096) or C3 (B=64). usage: TAG c0|c3
set -u
mkdir -p gpurun_out
TAG=${1:-c0}; W=${2:-c0}
if [ $W = c0 ]; then
CASE=c0 timeout 300 python scripts/c0_time.py 2>&1 | tail -1
CASE=c0 timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_eval|k_rows_n" -s 4 -c 1 -o gpurun_out/prof_${TAG}_c0fd -f python scripts/c0_time.py > gpurun_out/ncu_${TAG}_c0.log 2>&1
echo "ncu c0 rc=$?"
else
BATCH=64 JAC=fd timeout 300 python scripts/c3_once.py | tail -1
BATCH=64 JAC=fd timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_eval -s 4 -c 1 -o gpurun_out/prof_${TAG}_c3fd -f python scripts/c3_once.py > gpurun_out/ncu_${TAG}_c3.log 2>&1
echo "ncu c3 rc=$?"
fi
