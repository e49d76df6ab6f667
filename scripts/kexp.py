#!/usr/bin/env python
"""Kernel timing experiments on the bench workload (C2): which outputs cost what. GPU only."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from etol_b200 import capi, workloads as W

B = int(os.environ.get("KEXP_BATCH", "4096"))
wl = W.pm3d(batch=B)
ev = capi.Evaluator(wl, device=0)
dev = torch.device("cuda", 0)
x = torch.from_numpy(wl.x).to(dev)
f = torch.empty(B, dtype=torch.float64, device=dev)
g = torch.empty((B, ev.ncons), dtype=torch.float64, device=dev)
jac = torch.empty((B, ev.nnz), dtype=torch.float64, device=dev)
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)
st = torch.cuda.Stream(device=dev); torch.cuda.set_stream(st); sp = st.cuda_stream

FLUSH = os.environ.get("KEXP_FLUSH", "write")
rbuf = torch.zeros(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)

def timeit(fn, n=15):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for i in range(n):
        if FLUSH in ("write", "write+read"):
            flush.fill_(float(i))
        if FLUSH in ("read", "write+read"):
            rbuf.sum()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(st); fn(); e.record(st); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.median(ts)), float(np.min(ts))

res = {}
res["fill_446MB"] = timeit(lambda: jac.fill_(1.0))
res["copy_jac"] = timeit(lambda: jac.copy_(jac.roll(0)) if False else jac.mul_(1.0))
res["g_only"] = timeit(lambda: ev.eval_ptr(x.data_ptr(), f.data_ptr(), g.data_ptr(), None, capi.JAC_EXACT, capi.MEM_DEVICE, sp))
res["f_only"] = timeit(lambda: ev.eval_ptr(x.data_ptr(), f.data_ptr(), None, None, capi.JAC_EXACT, capi.MEM_DEVICE, sp))
res["jac_exact_only"] = timeit(lambda: ev.eval_ptr(x.data_ptr(), None, None, jac.data_ptr(), capi.JAC_EXACT, capi.MEM_DEVICE, sp))
res["jac_fd_only"] = timeit(lambda: ev.eval_ptr(x.data_ptr(), None, None, jac.data_ptr(), capi.JAC_FD, capi.MEM_DEVICE, sp))
res["all_exact"] = timeit(lambda: ev.eval_ptr(x.data_ptr(), f.data_ptr(), g.data_ptr(), jac.data_ptr(), capi.JAC_EXACT, capi.MEM_DEVICE, sp))
res["all_fd"] = timeit(lambda: ev.eval_ptr(x.data_ptr(), f.data_ptr(), g.data_ptr(), jac.data_ptr(), capi.JAC_FD, capi.MEM_DEVICE, sp))
print(FLUSH, json.dumps({k: [round(v[0], 4), round(v[1], 4)] for k, v in res.items()}))
