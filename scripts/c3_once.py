import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from etol_b200 import capi, workloads as W
B = int(os.environ.get("BATCH", "1")); mode = capi.JAC_FD if os.environ.get("JAC", "fd") == "fd" else capi.JAC_EXACT
wl = W.fw6(batch=B); ev = capi.Evaluator(wl, device=0); dev = torch.device("cuda", 0)
x = torch.from_numpy(wl.x).to(dev)
f = torch.empty(B, dtype=torch.float64, device=dev); g = torch.empty((B, ev.ncons), dtype=torch.float64, device=dev)
jac = torch.empty((B, ev.nnz), dtype=torch.float64, device=dev)
for _ in range(6):
    ev.eval_ptr(x.data_ptr(), f.data_ptr(), g.data_ptr(), jac.data_ptr(), mode, capi.MEM_DEVICE, None)
ev.sync(); torch.cuda.synchronize(); print("ok")
