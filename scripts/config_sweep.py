#!/usr/bin/env python
"""Evaluation rates of the other BASELINE.json configurations (parity-test cases, not bench lines):
device-resident kernel time per batch for both Jacobian modes, and the host-call latency of the
IPOPT-shaped single-instance shims. GPU only. Writes profiles-ready JSON to stdout."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from etol_b200 import capi, workloads as W

dev = torch.device("cuda", 0)
st = torch.cuda.Stream(device=dev); torch.cuda.set_stream(st); sp = st.cuda_stream
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)
sweep = torch.zeros(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)

CONFIGS = {
    "C0 reference VGP (si2d, 33 nodes, ocp_2d_ex1 shape), B=4096": lambda: W.reference_vgp("ocp", batch=4096, jitter=0.02),
    "C1 pm3d 40 nodes 8 cylinders, B=1": lambda: W.pm3d(batch=1),
    "C2 pm3d 40 nodes 8 cylinders, B=4096": lambda: W.pm3d(batch=4096),
    "C3 fw6 200 nodes 64 cylinders, B=1": lambda: W.fw6(batch=1),
    "C3 fw6 200 nodes 64 cylinders, B=64": lambda: W.fw6(batch=64),
    "C4 pm3d 3 phases x 30 nodes, B=1024": lambda: W.pm3d_multiphase(batch=1024),
}

def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for i in range(n):
        flush.fill_(float(i)); sweep.sum()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(st); fn(); e.record(st); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.median(ts))

out = {}
for name, mk in CONFIGS.items():
    wl = mk()
    ev = capi.Evaluator(wl, device=0)
    B = wl.batch
    x = torch.from_numpy(wl.x).to(dev)
    f = torch.empty(B, dtype=torch.float64, device=dev)
    g = torch.empty((B, ev.ncons), dtype=torch.float64, device=dev)
    jac = torch.empty((B, ev.nnz), dtype=torch.float64, device=dev)
    row = {"nvars": ev.nvars, "ncons": ev.ncons, "nnz": ev.nnz, "groups": ev.dims.ngroups, "batch": B}
    for mode, tag in ((capi.JAC_FD, "fd"), (capi.JAC_EXACT, "exact")):
        ms = timeit(lambda: ev.eval_ptr(x.data_ptr(), f.data_ptr(), g.data_ptr(), jac.data_ptr(), mode, capi.MEM_DEVICE, sp))
        row[f"{tag}_ms_per_batch"] = round(ms, 4)
        row[f"{tag}_evals_per_s"] = round(B / (ms * 1e-3), 1)
        row[f"{tag}_hbm_GBps"] = round(8 * (ev.nvars + 1 + ev.ncons + ev.nnz) * B / (ms * 1e-3) / 1e9, 1)
    if B == 1:  # latency of one IPOPT-shaped callback with host buffers (copies inside)
        xh = np.ascontiguousarray(wl.x[0]); gh = np.zeros(ev.ncons); vh = np.zeros(ev.nnz); obj = np.zeros(1)
        import ctypes as C
        dp = C.POINTER(C.c_double)
        L = ev.L
        def cb_g(): L.ecuda_ipopt_eval_g(ev.h, ev.nvars, xh.ctypes.data_as(dp), 1, ev.ncons, gh.ctypes.data_as(dp))
        def cb_j(): L.ecuda_ipopt_eval_jac_g(ev.h, ev.nvars, xh.ctypes.data_as(dp), 1, ev.ncons, ev.nnz, None, None, vh.ctypes.data_as(dp))
        for fn, tag in ((cb_g, "ipopt_eval_g_us"), (cb_j, "ipopt_eval_jac_g_us")):
            for _ in range(20): fn()
            t0 = time.perf_counter()
            for _ in range(200): fn()
            row[tag] = round((time.perf_counter() - t0) / 200 * 1e6, 1)
    out[name] = row
    ev.close()
print(json.dumps(out, indent=1))
