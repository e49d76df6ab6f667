#!/bin/bash
# user models on the GPU: parity tests, then kernel time of pm3d-as-user-model next to the built-in pm3d
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "user" 2>&1 | tail -15
timeout 600 python scripts/user_model_time.py 2>&1 | tail -8
