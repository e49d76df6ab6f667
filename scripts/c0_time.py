"""C0 (reference VGP shape, si2d, 33 nodes) kernel time per batch of 4096, both Jacobian modes"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from etol_b200 import capi, workloads as W
dev = torch.device("cuda", 0)
B = 4096
which = os.environ.get("CASE", "c0")
wl = W.reference_vgp("ocp", batch=B, jitter=0.02) if which == "c0" else W.reference_vgp("mip", batch=B, jitter=0.02) if which == "c0mip" else W.fw6(batch=int(os.environ.get("BATCH", "64"))) if which == "c3" else W.pm3d_multiphase(batch=1024)
B = wl.batch
ev = capi.Evaluator(wl, device=0)
x = torch.from_numpy(wl.x).to(dev)
f = torch.empty(B, dtype=torch.float64, device=dev); g = torch.empty((B, ev.ncons), dtype=torch.float64, device=dev)
jac = torch.empty((B, ev.nnz), dtype=torch.float64, device=dev)
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); st = stream.cuda_stream
res = {}
for mode, tag in ((capi.JAC_FD, "fd_ms"), (capi.JAC_EXACT, "exact_ms")):
    fn = lambda: ev.eval_ptr(x.data_ptr(), f.data_ptr(), g.data_ptr(), jac.data_ptr(), mode, capi.MEM_DEVICE, st)
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for i in range(15):
        flush.fill_(float(i)); flush.sum()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(stream); fn(); e.record(stream); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    res[tag] = round(float(np.median(ts)), 4)
print(which, {k: os.environ[k] for k in os.environ if k.startswith("ECUDA_")}, res)
