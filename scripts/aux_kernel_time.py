"""Kernel time per batch of the Hessian, discretisation-error and resampling kernels (C2 shape, device buffers)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch
from etol_b200 import capi, workloads as W
dev = torch.device("cuda", 0)
B = int(os.environ.get("BATCH", "4096"))
wl = W.pm3d(batch=B)
ev = capi.Evaluator(wl, device=0)
L = capi.lib()
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); st = stream.cuda_stream
x = torch.from_numpy(wl.x).to(dev)
lam = torch.randn((B, ev.ncons), dtype=torch.float64, device=dev)
n = C.c_int32(0); L.ecuda_get_hess_structure(ev.h, C.byref(n), None, None)
hv = torch.empty((B, n.value), dtype=torch.float64, device=dev)
err = torch.empty((B, wl.nnodes[0] - 1), dtype=torch.float64, device=dev)
nn = np.array([61], dtype=np.int32); nv = (wl.ns + wl.nc) * 61 + 2
xn = torch.empty((B, nv), dtype=torch.float64, device=dev)
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)
def t(fn, k=15):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for i in range(k):
        flush.fill_(float(i)); flush.sum()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(stream); fn(); e.record(stream); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    return float(np.median(ts))
out = {"batch": B, "nnz_h": n.value}
ms = t(lambda: L.ecuda_eval_hess(ev.h, x.data_ptr(), None, 1.0, lam.data_ptr(), hv.data_ptr(), capi.MEM_DEVICE, st))
out["k_hess"] = {"ms": round(ms, 4), "evals_per_s": round(B / ms * 1e3), "GBps": round(8.0 * B * (ev.nvars + ev.ncons + n.value) / ms / 1e6, 1)}
ms = t(lambda: L.ecuda_ode_error(ev.h, x.data_ptr(), err.data_ptr(), capi.MEM_DEVICE, st))
out["k_ode_error"] = {"ms": round(ms, 4), "instances_per_s": round(B / ms * 1e3)}
ms = t(lambda: L.ecuda_resample(ev.h, x.data_ptr(), nn.ctypes.data_as(capi._ip), None, xn.data_ptr(), capi.MEM_DEVICE, st))
out["k_resample_40_to_61"] = {"ms": round(ms, 4), "instances_per_s": round(B / ms * 1e3)}
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True); json.dump(out, open("gpurun_out/aux_kernel_time.json", "w"), indent=1)
