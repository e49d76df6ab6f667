"""fused summary (+ all-gather epilogue, 1 rank) next to the plain kernel; compact vs dense bounds (ECUDA_DENSE_BOUNDS=1)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from etol_b200 import capi, workloads as W
dev = torch.device("cuda", 0)
wl = W.pm3d(batch=4096)
ev = capi.Evaluator(wl, device=0)
B = wl.batch
x = torch.from_numpy(wl.x).to(dev)
f = torch.empty(B, dtype=torch.float64, device=dev); g = torch.empty((B, ev.ncons), dtype=torch.float64, device=dev)
jac = torch.empty((B, ev.nnz), dtype=torch.float64, device=dev)
buf = torch.empty((B, 2), dtype=torch.float64, device=dev); summ = torch.empty((B, 2), dtype=torch.float64, device=dev)
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); st = stream.cuda_stream
def t(fn, n=15):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for i in range(n):
        flush.fill_(float(i)); flush.sum()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(stream); fn(); e.record(stream); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    return round(float(np.median(ts)), 4)
out = {}
for mode, tag in ((capi.JAC_FD, "fd"), (capi.JAC_EXACT, "exact")):
    plain = lambda: ev.eval_ptr(x.data_ptr(), f.data_ptr(), g.data_ptr(), jac.data_ptr(), mode, capi.MEM_DEVICE, st)
    fused = lambda: ev.eval_allgather_ptr(x.data_ptr(), f.data_ptr(), g.data_ptr(), jac.data_ptr(), mode, [buf.data_ptr()], 0, st)
    out[tag] = {"plain": t(plain), "fused": t(fused)}
    buf.zero_(); fused(); ev.summarize_ptr(f.data_ptr(), g.data_ptr(), summ.data_ptr(), st); torch.cuda.synchronize()
    out[tag]["fused == k_summary"] = bool(torch.equal(buf, summ))
print("dense bounds" if os.environ.get("ECUDA_DENSE_BOUNDS") else "compact bounds", out)
