"""Kernel time of a model compiled at run time (pm3d recorded as callbacks) next to the built-in pm3d, C2 shape."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from etol_b200 import capi, workloads as W

dev = torch.device("cuda", 0)
B = int(os.environ.get("BATCH", "4096"))
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); st = stream.cuda_stream
out = {}
for name, mk in (("builtin pm3d", lambda: W.pm3d(batch=B)), ("user pm3d", lambda: W.pm3d_user(batch=B)),
                 ("user unicycle N33", lambda: W.unicycle(batch=B, ntracks=1)), ("user dragmass N33", lambda: W.dragmass(batch=B))):
    wl = mk()
    t0 = time.time(); ev = capi.Evaluator(wl, device=0); setup = time.time() - t0
    x = torch.from_numpy(wl.x).to(dev)
    f = torch.empty(B, dtype=torch.float64, device=dev); g = torch.empty((B, ev.ncons), dtype=torch.float64, device=dev)
    jac = torch.empty((B, ev.nnz), dtype=torch.float64, device=dev)
    res = {"setup_s": round(setup, 2)}
    for mode, tag in ((capi.JAC_FD, "fd_ms"), (capi.JAC_EXACT, "exact_ms")):
        fn = lambda: ev.eval_ptr(x.data_ptr(), f.data_ptr(), g.data_ptr(), jac.data_ptr(), mode, capi.MEM_DEVICE, st)
        for _ in range(3): fn()
        torch.cuda.synchronize(); ts = []
        for i in range(15):
            flush.fill_(float(i)); flush.sum()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(stream); fn(); e.record(stream); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
        res[tag] = round(float(np.median(ts)), 4)
        res[tag.replace("_ms", "_GBps")] = round(8.0 * B * (ev.nvars + 1 + ev.ncons + ev.nnz) / (res[tag] * 1e-3) / 1e9, 1)
    out[name] = res
    ev.close()
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/user_model_time.json", "w"), indent=1)
