"""latency of one C3 evaluation (fw6, 200 nodes, 64 cylinders) against the number of CTA slices per instance"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1:  # child: one setting
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np, torch
    from etol_b200 import capi, workloads as W
    import oracle_binding as ob
    B = int(sys.argv[1])
    wl = W.fw6(batch=B)
    ev = capi.Evaluator(wl, device=0)
    dev = torch.device("cuda", 0)
    x = torch.from_numpy(wl.x).to(dev)
    f = torch.empty(B, dtype=torch.float64, device=dev); g = torch.empty((B, ev.ncons), dtype=torch.float64, device=dev)
    jac = torch.empty((B, ev.nnz), dtype=torch.float64, device=dev)
    stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); st = stream.cuda_stream
    res = {}
    for mode, tag in ((capi.JAC_FD, "fd_ms"), (capi.JAC_EXACT, "exact_ms")):
        fn = lambda: ev.eval_ptr(x.data_ptr(), f.data_ptr(), g.data_ptr(), jac.data_ptr(), mode, capi.MEM_DEVICE, st)
        for _ in range(3): fn()
        torch.cuda.synchronize(); ts = []
        for i in range(20):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(stream); fn(); e.record(stream); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
        res[tag] = round(float(np.median(ts)), 4)
    if B == 1:  # parity of this slicing against the oracle (FD bit-identical)
        got = ev.eval_host(wl.x, jac_mode=capi.JAC_FD)
        ref = ob.Oracle(wl).eval(wl.x, jac_mode=capi.JAC_FD, style=1, nthreads=8)
        res["fd_bitwise"] = bool(np.array_equal(got["jac"], ref["jac"]) and np.array_equal(got["g"], ref["g"]))
    print(json.dumps(res))
else:
    out = {}
    for B in (1, 8, 64):
        for sl in ("5", "10", "25", "50", "auto"):
            env = dict(os.environ)
            if sl != "auto": env["ECUDA_SLICES"] = sl
            else: env.pop("ECUDA_SLICES", None)
            r = subprocess.run([sys.executable, __file__, str(B)], env=env, capture_output=True, text=True, timeout=600)
            out[f"B={B} slices={sl}"] = json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else r.stderr[-300:]
            print(f"B={B} slices={sl}", out[f"B={B} slices={sl}"], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "slice_sweep.json"), "w"), indent=1)
