#!/bin/bash
# plugin on the GPU: plugin tests (built-in and user models), the three example executables
set -u
timeout 900 python -m pytest tests/test_plugin.py -m gpu -x -q 2>&1 | tail -12
timeout 600 python scripts/run_examples.py 2>&1 | tail -40
