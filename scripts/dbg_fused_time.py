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
flags = torch.zeros(64, dtype=torch.int64, device=dev)
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev); sweep = torch.zeros_like(flush)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); st = stream.cuda_stream
step = [0]
def t(fn, n=15):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for i in range(n):
        flush.fill_(float(i)); sweep.sum()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(stream); fn(); e.record(stream); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    return round(float(np.median(ts)), 4)
def plain(): ev.eval_ptr(x.data_ptr(), f.data_ptr(), g.data_ptr(), jac.data_ptr(), capi.JAC_FD, capi.MEM_DEVICE, st)
def fused(): ev.eval_allgather_ptr(x.data_ptr(), f.data_ptr(), g.data_ptr(), jac.data_ptr(), capi.JAC_FD, [buf.data_ptr()], 0, st)
def fused_bar():
    fused(); step[0] += 1; ev.peer_barrier_ptr([flags.data_ptr()], 0, step[0], st)
def sep(): plain(); ev.summarize_ptr(f.data_ptr(), g.data_ptr(), summ.data_ptr(), st)
print({"plain": t(plain), "fused": t(fused), "fused+barrier(1 rank)": t(fused_bar), "plain+k_summary": t(sep)})
fused(); sep(); torch.cuda.synchronize()
print("fused == k_summary:", bool(torch.equal(buf, summ)), float(summ[:, 1].max()))
def fused_ex(): ev.eval_allgather_ptr(x.data_ptr(), f.data_ptr(), g.data_ptr(), jac.data_ptr(), capi.JAC_EXACT, [buf.data_ptr()], 0, st)
def plain_ex(): ev.eval_ptr(x.data_ptr(), f.data_ptr(), g.data_ptr(), jac.data_ptr(), capi.JAC_EXACT, capi.MEM_DEVICE, st)
print({"plain exact": t(plain_ex), "fused exact": t(fused_ex)})
buf.zero_(); fused_ex(); torch.cuda.synchronize()
print("fused exact == k_summary:", bool(torch.equal(buf, summ)))
