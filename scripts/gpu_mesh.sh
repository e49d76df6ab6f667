#!/bin/bash
set -u
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ode_error or resample or round_trip or hessian" 2>&1 | tail -15
timeout 900 python -m pytest tests/test_plugin.py -m gpu -x -q 2>&1 | tail -12
timeout 600 python scripts/run_examples.py 2>&1 | grep -A8 "example3"
