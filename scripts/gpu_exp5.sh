#!/bin/bash
set -u
for L in libecuda.so libecuda_rows4.so; do echo $L; ECUDA_LIB=$PWD/etol_b200/csrc/$L KEXP_FLUSH=write+read timeout 300 python scripts/kexp.py 2>&1 | tail -1; done
