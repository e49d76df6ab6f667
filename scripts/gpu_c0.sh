#!/bin/bash
# C0 / C4 kernel times + GPU parity tests
set -u
mkdir -p gpurun_out
CASE=c0 timeout 300 python scripts/c0_time.py 2>&1 | tail -1
CASE=c4 timeout 300 python scripts/c0_time.py 2>&1 | tail -1
if [ "${1:-}" = "tests" ]; then timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4; fi
