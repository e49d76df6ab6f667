#!/usr/bin/env python
"""Small evaluations through every kernel family, meant to run under compute-sanitizer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from etol_b200 import capi, workloads as W
cases = [W.pm3d(batch=3), W.reference_vgp("ocp", batch=2, jitter=0.01), W.pm3d_multiphase(batch=2, scaled=True),
         W.fw6(batch=1, nnodes=41, ncyl=4), W.pm3d(batch=2, nnodes=9, ncyl=0)]
if os.environ.get("SANITIZE_USER", "1") == "1":  # round 2: run-time compiled models (time-dependent, traced path rows)
    cases += [W.zone(batch=2, ntracks=1, timedep=True), W.gust(batch=2, nnodes=12, ncyl=2)]
for wl in cases:
    user = getattr(wl, "tape", None) is not None
    for env in ({}, {"ECUDA_IMAGE": "1"}, {"ECUDA_NO_COPY_WARP": "1"}, {"ECUDA_NO_FAST": "1"}, {"ECUDA_NO_ROWSN": "1"},
                {"ECUDA_EXACT_KERNEL": "ring"}, {"ECUDA_EXACT_KERNEL": "stream"}, {"ECUDA_HOST_CHUNKS": "2"}):
        if user and env and "ECUDA_NO_ROWSN" not in env and "ECUDA_NO_FAST" not in env:
            continue
        os.environ.update(env)
        ev = capi.Evaluator(wl, device=0)
        for k in env: os.environ.pop(k)
        for mode in (capi.JAC_FD, capi.JAC_EXACT):
            r = ev.eval_host(wl.x, want=("f", "g", "jac", "grad"), jac_mode=mode)
            assert np.isfinite(r["jac"]).all() and np.isfinite(r["g"]).all()
        if not user:  # compact exact Jacobian (k_gather_local) through host buffers
            idx, shared = ev.compact_structure()
            c = ev.eval_compact_host(wl.x, idx.size)
            assert np.array_equal(ev.splice(shared, idx, c["jac_local"]), r["jac"])
        if wl.gl is not None:
            ev.summary_host(wl.x)
        ev.close()
print("sanitize workload done")
