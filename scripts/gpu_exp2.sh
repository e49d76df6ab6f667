#!/bin/bash
set -u
for F in write write+read read none; do KEXP_FLUSH=$F timeout 300 python scripts/kexp.py 2>&1 | tail -1; done
