#!/usr/bin/env python
"""bench.py -- NLP constraint+Jacobian evaluations per second over batched VGPs (BASELINE.json).

One "step" = one pass of the hot path over one batch: for every instance of the batch the
objective f, the constraint vector g[ncons] and the sparse Jacobian values J[nnz] (index-set
finite differences, triplet layout) at one fixed decision vector.

  python bench.py [--gpus N] [--steps K] [--warmup W]          the CUDA path (this repo)
  python bench.py --impl reference [...]                       the CPU path on the host cores

Workload at N=1: BASELINE.json configs[1] (C2) -- the 3-D point-mass UAS VGP (8 cylinders, 40 LGL nodes)
batched to 4096 random instances. N>1 (torchrun, one process per GPU): BASELINE.json configs[4] (C5) -- the
scenario sweep of 65 536 instances of the same VGP, sharded in contiguous ranges of 65536/N instances per GPU;
every step the ranks exchange per-instance summaries {f, max violation} (fused into the evaluation kernel, P2P
stores over NVLink) and, measured separately on the same line ("gather_full"), all-gather the full per-instance
results [g | Jvals] with NCCL.

The `--impl reference` arm times the reference's CPU algorithm for the same path. The reference's
own binaries (PSOPT 5.0.0 + ADOL-C + IPOPT) cannot be built in this image,
so it runs the oracle port: per-node std::function/std::any callbacks like ePSOPT::dae and the
column-grouped finite-difference Jacobian, on all host threads (cpu_baseline.kind = "port").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# rank 0 prints exactly one line on stdout: keep NCCL's version banner (NCCL_DEBUG=VERSION) off it
if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
    os.environ["NCCL_DEBUG"] = "WARN"

import numpy as np  # noqa: E402

METRIC = "nlp_constraint_jacobian_evals_per_sec"
UNIT = "evals/s"
BATCH_PER_GPU = 4096   # C2, the N=1 workload
C5_INSTANCES = 65536   # C5, sharded over the ranks when N > 1


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ecuda", choices=["ecuda", "reference"])
    ap.add_argument("--batch", type=int, default=0,
                    help="instances per GPU (default: 4096 at N=1 = C2; 65536/N at N>1 = C5)")
    ap.add_argument("--jac", default="fd", choices=["fd", "exact"])
    ap.add_argument("--gather", default="summary", choices=["summary", "full", "none"])
    ap.add_argument("--nccl-gather", action="store_true",
                    help="exchange the per-instance summaries with NCCL all_gather instead of the fused P2P kernel")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the exact-mode line, the other BASELINE configurations, the D2H probe and gather_full")
    args = ap.parse_args()
    if args.batch <= 0:
        args.batch = BATCH_PER_GPU if args.gpus <= 1 else C5_INSTANCES // args.gpus
    return args


def workload_config(wl, args, n_gpus):
    """identical in both arms (the driver compares the dicts)"""
    total = wl.batch * n_gpus
    name = ("C2: pm3d UAS VGP (6 states, 3 controls, 8 cylinders, 40 Legendre nodes), 4096 random instances"
            if n_gpus == 1 and wl.batch == BATCH_PER_GPU else
            f"C5: scenario sweep of {total} pm3d UAS VGP instances (6 states, 3 controls, 8 cylinders, 40 Legendre "
            f"nodes) sharded over {n_gpus} GPU(s) in contiguous ranges of {wl.batch}" if total == C5_INSTANCES else
            f"pm3d UAS VGP (C2 shape), {wl.batch} random instances per GPU on {n_gpus} GPU(s)")
    return {"workload": name + ", evaluation only at fixed decision vectors",
            "instances_per_gpu": wl.batch, "n_instances": total, "nvars": wl.nvars, "ncons": wl.ncons,
            "jacobian": "fd_indexset" if args.jac == "fd" else "exact", "pattern": "dense_node",
            "l2": "flushed between timed steps: 256 MiB write, then a 256 MiB read sweep so the flush buffer's "
                  "dirty lines are written back before the timed region; outputs per step exceed L2",
            "parallelism": f"instances sharded over {n_gpus} GPU(s), gather={args.gather if n_gpus > 1 else 'n/a'}",
            "timing": "before every timed step the ranks are aligned by an untimed in-stream barrier, so that the skew "
                      "of the untimed L2-flush kernels is not charged to the step's exchange" if n_gpus > 1 else
                      "CUDA events on the launching stream around each step"}


def shard_workload(args, world, rank):
    """N=1: C2 (4096 instances, seed W.SEED). N>1: rank r owns the contiguous range [r*B, (r+1)*B) of the sweep."""
    from etol_b200 import shard, workloads as W
    if world == 1:
        return W.pm3d(batch=args.batch)
    full = W.pm3d(batch=args.batch * world)
    lo, hi = shard.shard_range(full.batch, world, rank)  # [r*B/G, (r+1)*B/G)
    return full.slice_batch(lo, hi)



# ---- clocks sampler (B200_PROFILING.md recipe) -----------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t_begin, t_end):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                clk, mx = float(parts[1]), float(parts[2])
            except ValueError:
                continue
            smax = mx
            if t_begin - 0.05 <= ts <= t_end + 0.15:
                sm.append(clk)
                for name, val in zip(names, parts[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        if not sm:  # region shorter than the sampling period: use every sample taken
            for ts, line in self.rows:
                parts = [p.strip() for p in line.split(",")]
                try:
                    sm.append(float(parts[1]))
                except (ValueError, IndexError):
                    pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---- the CPU arm ---------------------------------------------------------------------------------------------
def cpu_sample(wl, seconds_budget, jac_mode, style=0):
    """times the oracle on a bounded sample of the workload, all host threads"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding as ob
    # all host threads this process may use; torchrun exports OMP_NUM_THREADS=1, which must not
    # shrink the CPU arm (the oracle takes its thread count explicitly)
    try:
        nthr = len(os.sched_getaffinity(0))
    except AttributeError:
        nthr = os.cpu_count() or 1
    probe_n = min(wl.batch, nthr)
    sub = wl.slice_batch(0, min(wl.batch, 64 * nthr))
    orc = ob.Oracle(sub)
    r = orc.eval(sub.x[:probe_n], want=("f", "g", "jac"), jac_mode=jac_mode, style=style, nthreads=nthr,
                 count=probe_n)
    per_round = max(r["seconds"], 1e-6)
    rounds = int(max(1, min(seconds_budget / per_round, sub.batch // probe_n)))
    n = probe_n * rounds
    r = orc.eval(sub.x[:n], want=("f", "g", "jac"), jac_mode=jac_mode, style=style, nthreads=nthr, count=n)
    return {"value": n / r["seconds"], "unit": UNIT, "cores": nthr, "kind": "port",
            "sample": f"{n} of {wl.batch} instances, f+g+J({'fd_indexset' if jac_mode == 1 else 'exact'}), "
                      f"{'reference-style std::function/std::any callbacks' if style == 0 else 'tight loops'}, "
                      f"OpenMP over instances, {r['seconds']:.2f} s",
            "seconds": r["seconds"], "n": n}


def cpu_rowrestricted_sample(wl, seconds_budget):
    """tight loops + row-restricted finite differences on all host threads: the kernel logic of etol_b200/csrc stepped
    on the CPU (tests/emu, the N-specialised row-owner variant), one instance range per thread"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import emu_binding as eb
    from concurrent.futures import ThreadPoolExecutor
    try:
        nthr = len(os.sched_getaffinity(0))
    except AttributeError:
        nthr = os.cpu_count() or 1
    eb.lib()
    per = 4

    def work(i):
        sub = wl.slice_batch(i * per, (i + 1) * per)
        eb.emu_eval(sub, sub.x, want=("f", "g", "jac"), jac_mode=1, nthr=256, variant="rowsn")
        return per

    work(0)
    nchunks = wl.batch // per
    n, t0 = 0, time.time()
    with ThreadPoolExecutor(max_workers=nthr) as pool:
        while time.time() - t0 < seconds_budget:  # whole passes over the batch until the budget is used
            n += sum(pool.map(lambda i: work(i % nchunks), range(nchunks)))
    dt = time.time() - t0
    return {"value": n / dt, "sample": f"{n} of {wl.batch} instances, f+g+J(fd_indexset, row-restricted), kernel logic "
                                        f"stepped on the CPU, {nthr} threads, {dt:.2f} s"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = shard_workload(args, max(1, args.gpus), 0)
    jac_mode = 1 if args.jac == "fd" else 0
    per_step_budget = max(0.5, min(8.0, 120.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        cpu_sample(wl, per_step_budget / 4, jac_mode)
    tot_n, tot_s, last = 0, 0.0, None
    for _ in range(args.steps):
        last = cpu_sample(wl, per_step_budget, jac_mode)
        tot_n += last["n"]
        tot_s += last["seconds"]
    value = tot_n / tot_s
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / args.steps,
            "higher_is_better": True, "scaling": "weak" if args.gpus <= 1 else "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(wl, args, max(1, args.gpus)),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": last["cores"], "kind": "port",
                             "sample": f"each step: {last['sample']}"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "oracle port of the reference CPU path (PSOPT/ADOL-C cannot be built here); host cores only"}
    emit(line)
    return 0


# ---- the CUDA arm ----------------------------------------------------------------------------------------------
def run_ecuda(args):
    import torch
    import torch.distributed as dist
    from etol_b200 import capi, workloads as W

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the eCUDA path has no CPU fallback (use --impl reference "
                         "for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    jac_mode = capi.JAC_FD if args.jac == "fd" else capi.JAC_EXACT
    # N=1: C2. N>1: this rank's contiguous range of the C5 sweep
    wl = shard_workload(args, world, rank)
    ev = capi.Evaluator(wl, device=local)
    B, nv, ng, nz = wl.batch, ev.nvars, ev.ncons, ev.nnz
    torch.cuda.synchronize()
    x = torch.from_numpy(wl.x).to(dev)
    f = torch.empty(B, dtype=torch.float64, device=dev)
    g = torch.empty((B, ng), dtype=torch.float64, device=dev)
    jac = torch.empty((B, nz), dtype=torch.float64, device=dev)
    summ = torch.empty((B, 2), dtype=torch.float64, device=dev)
    gathered = torch.empty((world * B, 2), dtype=torch.float64, device=dev) if world > 1 else None
    full_gather = None
    if world > 1 and args.gather == "full":
        full_gather = torch.empty((world * B, nz), dtype=torch.float64, device=dev)
    # fused summary + all-gather over NVLink peer memory (ecuda_summarize_allgather): two symmetric
    # buffers used alternately, so a rank that runs one step ahead never overwrites rows a peer may
    # still be reading; falls back to NCCL all_gather when symmetric memory is unavailable
    peers, hdls, gather_impl = None, None, "n/a"
    if world > 1 and args.gather != "none":
        gather_impl = "nccl all_gather_into_tensor"
        if not args.nccl_gather:
            try:
                import torch.distributed._symmetric_memory as symm
                bufs = [symm.empty((world * B, 2), dtype=torch.float64, device=dev) for _ in range(2)]
                hdls = [symm.rendezvous(t, dist.group.WORLD) for t in bufs]
                peers = [[int(p) for p in hd.buffer_ptrs] for hd in hdls]
                sym_bufs = bufs
                # counters of ecuda_peer_barrier (one array of `world` 64-bit slots per rank)
                flagbuf = symm.empty(64, dtype=torch.int64, device=dev)
                flagbuf.zero_()
                flag_hdl = symm.rendezvous(flagbuf, dist.group.WORLD)
                flag_ptrs = [int(p) for p in flag_hdl.buffer_ptrs]
                torch.cuda.synchronize()
                flag_hdl.barrier(channel=0)  # every rank has zeroed its counters before anyone bumps them
                gather_impl = ("fused into k_eval: CTA epilogue stores {f, max violation} to every rank over NVLink "
                               "(symmetric memory), 1 flag barrier per step (ecuda_peer_barrier)")
            except Exception as exc:  # noqa: BLE001
                peers, hdls = None, None
                gather_impl = f"nccl all_gather_into_tensor (symmetric memory unavailable: {type(exc).__name__})"
    step_no = [0]
    gather_parity = None
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)
    sweep = torch.zeros(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)

    def flush_l2(i):
        # untimed. The write evicts everything; the read sweep then evicts the write's dirty lines, so
        # their write-back does not land inside the timed step (it costs a store-bound kernel ~25 %).
        flush.fill_(float(i))
        sweep.sum()
    # a dedicated (non-default) stream: kernels, copies, NCCL and the timing events all go on it
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    sp = stream.cuda_stream
    assert sp != 0

    def step():
        if world > 1 and args.gather != "none" and peers is not None:
            # ONE kernel: evaluation, and every CTA's epilogue stores its instance's {f, max violation}
            # row into all ranks' gathered buffers over NVLink; then one cross-GPU barrier
            s = step_no[0] & 1
            step_no[0] += 1
            ev.eval_allgather_ptr(x.data_ptr(), f.data_ptr(), g.data_ptr(), jac.data_ptr(), jac_mode, peers[s], rank, sp)
            ev.peer_barrier_ptr(flag_ptrs, rank, step_no[0], sp)
            if full_gather is not None:
                dist.all_gather_into_tensor(full_gather, jac)
            return
        ev.eval_ptr(x.data_ptr(), f.data_ptr(), g.data_ptr(), jac.data_ptr(), jac_mode, capi.MEM_DEVICE, sp)
        if world > 1 and args.gather != "none":
            # per-instance {f, max violation} from the f, g just computed, gathered on every rank
            if False:
                pass
            else:
                ev.summarize_ptr(f.data_ptr(), g.data_ptr(), summ.data_ptr(), sp)
                dist.all_gather_into_tensor(gathered, summ)
            if full_gather is not None:
                dist.all_gather_into_tensor(full_gather, jac)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    if peers is not None:  # the fused exchange must give what NCCL gives
        ev.summarize_ptr(f.data_ptr(), g.data_ptr(), summ.data_ptr(), sp)
        dist.all_gather_into_tensor(gathered, summ)
        torch.cuda.synchronize()
        last = sym_bufs[(step_no[0] - 1) & 1]
        if not torch.equal(last, gathered):
            raise SystemExit("bench.py: fused P2P all-gather disagrees with NCCL all_gather")
        gather_parity = "bit-equal (fused P2P exchange == summary kernel + NCCL all_gather, all rows, this run)"
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    launches0 = ev.launch_count()
    barrier()
    t_begin = time.time()
    def align():
        # untimed, in-stream: all ranks leave their L2 flush before anyone starts the timed step, so the
        # step's exchange does not absorb the skew of the flush kernels
        if world > 1:
            if hdls is not None:
                hdls[0].barrier(channel=1)
            else:
                dist.all_reduce(skew_token)

    skew_token = torch.zeros(1, device=dev)
    for i in range(args.steps):
        flush_l2(i)                    # untimed: evict L2 between timed steps
        align()
        starts[i].record(stream)
        step()
        stops[i].record(stream)
    barrier()
    t_end = time.time()
    launches = ev.launch_count() - launches0
    per_step_ms = [s.elapsed_time(e) for s, e in zip(starts, stops)]
    total_ms = float(sum(per_step_ms))
    clocks = sampler.stop(t_begin, t_end)
    tt = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms_max = float(tt.item())
    value = world * B * args.steps / (total_ms_max / 1e3)

    # ---- roofline of the dominant kernel (k_eval): kernel-only timing on the same stream ------------
    kern_ms = []
    for i in range(args.steps):
        flush_l2(i)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(stream)
        ev.eval_ptr(x.data_ptr(), f.data_ptr(), g.data_ptr(), jac.data_ptr(), jac_mode, capi.MEM_DEVICE, sp)
        e.record(stream)
        torch.cuda.synchronize()
        kern_ms.append(s.elapsed_time(e))
    kern_avg_ms = float(np.mean(kern_ms))
    inst_bytes = 8 * ev.dims.inst_stride
    alg_bytes_unit = 8 * (nv + 1 + ng + nz) + inst_bytes
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"
    achieved = alg_bytes_unit * B / (kern_avg_ms / 1e3) / 1e9
    traffic, fp64, tsrc = None, None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and B == BATCH_PER_GPU:  # the ncu captures were taken at the default batch
        try:
            tj = json.load(open(tpath))
            traffic = tj.get(f"k_eval_{args.jac}_C2_bytes_per_launch")
            tsrc = tj.get(f"k_eval_{args.jac}_C2_source", "ncu capture") + " -- an ncu capture recorded in profiles/traffic.json, not a measurement of this run"
            ops = tj.get(f"k_eval_{args.jac}_C2_fp64_thread_instr_per_launch")
            if ops:
                # second roof of the finite-difference kernel: FP64 pipe. Peak measured now with a
                # register-only DFMA microbenchmark; executed FP64 thread instructions from ncu.
                peak_tf = ev.fp64_peak_tflops()
                lane_instr = ops["dfma"] + ops["dadd"] + ops["dmul"]
                fp64 = {"peak_tflops_measured": peak_tf, "thread_instr_per_launch": lane_instr,
                        "executed_tflops": (2 * ops["dfma"] + ops["dadd"] + ops["dmul"]) / (kern_avg_ms / 1e3) / 1e12,
                        "pipe_frac": lane_instr / (kern_avg_ms / 1e3) / (peak_tf * 1e12 / 2.0),
                        "note": "row-restricted FD needs ~21.5 FP64 instructions per D-coupled triplet; "
                                "pipe_frac = executed FP64 thread instructions / (time x measured DFMA issue rate)"}
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": tsrc, "kernel": "k_rows_n<pm3d,40,FD>" if args.jac == "fd" else "k_eval_rows<pm3d,5,exact>",
                "kernel_ms": kern_avg_ms, "algorithmic_bytes_per_unit": alg_bytes_unit, "units_per_launch": B,
                "peak_source": peak_src}
    if fp64:
        roofline["fp64"] = fp64

    def kernel_only(mode, n):
        ts = []
        for i in range(n):
            flush_l2(i)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(stream)
            ev.eval_ptr(x.data_ptr(), f.data_ptr(), g.data_ptr(), jac.data_ptr(), mode, capi.MEM_DEVICE, sp)
            e.record(stream)
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        return float(np.mean(ts))

    # ---- the reference's default derivative mode (derivatives = "automatic", ePSOPT.cpp:64): exact Jacobian,
    # same workload, same flush, kernel-only like the roofline above
    extras = {}
    if not args.no_extras:
        other = capi.JAC_EXACT if args.jac == "fd" else capi.JAC_FD
        oms = kernel_only(other, max(5, min(args.steps, 20)))
        och = alg_bytes_unit * B / (oms / 1e3) / 1e9
        extras["exact" if args.jac == "fd" else "fd"] = {
            "kernel_ms": oms, "achieved": och, "unit": "GB/s", "frac": och / peak, "evals_per_s": B / (oms / 1e3),
            "what": "same workload and flush, Jacobian mode " + ("exact (the reference's default)" if args.jac == "fd" else "fd")}
        # bare pinned device -> host copy of one step's Jacobian values on the same stream: the roof of the e2e number
        hprobe = torch.empty((B, nz), dtype=torch.float64).pin_memory()
        for _ in range(2):
            hprobe.copy_(jac, non_blocking=True)
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(stream)
        for _ in range(3):
            hprobe.copy_(jac, non_blocking=True)
        e.record(stream)
        barrier()
        tp = torch.tensor([s.elapsed_time(e) / 3.0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tp, op=dist.ReduceOp.MAX)
        extras["d2h_probe"] = {"bytes": int(8 * B * nz), "ms": float(tp.item()),
                               "GBps_per_gpu": 8 * B * nz / (float(tp.item()) / 1e3) / 1e9,
                               "what": "cudaMemcpyAsync device -> pinned host of one step's Jacobian values, all ranks "
                                       "at once, max over ranks"}
        del hprobe

    # ---- the other BASELINE configurations (parity-test cases): kernel time per batch in the same run, same clocks ----
    if world == 1 and not args.no_extras:
        others = {}
        cfgs = (("C0 reference VGP (si2d, 33 nodes, ocp_2d_ex1 shape), B=4096", lambda: W.reference_vgp("ocp", batch=4096, jitter=0.02)),
                ("C3 fw6 200 nodes 64 cylinders, B=64", lambda: W.fw6(batch=64)),
                ("C4 pm3d 3 phases x 30 nodes, B=1024", lambda: W.pm3d_multiphase(batch=1024)))
        for name, mk in cfgs:
            try:
                w2 = mk()
                e2 = capi.Evaluator(w2, device=local)
                x2 = torch.from_numpy(w2.x).to(dev)
                f2 = torch.empty(w2.batch, dtype=torch.float64, device=dev)
                g2 = torch.empty((w2.batch, e2.ncons), dtype=torch.float64, device=dev)
                j2 = torch.empty((w2.batch, e2.nnz), dtype=torch.float64, device=dev)
                unit = 8 * (e2.nvars + 1 + e2.ncons + e2.nnz) + 8 * e2.dims.inst_stride
                row = {"nvars": e2.nvars, "ncons": e2.ncons, "nnz": e2.nnz, "batch": w2.batch,
                       "algorithmic_bytes_per_unit": unit}
                for mode, tag in ((capi.JAC_FD, "fd"), (capi.JAC_EXACT, "exact")):
                    for _ in range(2):
                        e2.eval_ptr(x2.data_ptr(), f2.data_ptr(), g2.data_ptr(), j2.data_ptr(), mode, capi.MEM_DEVICE, sp)
                    ts = []
                    for i in range(5):
                        flush_l2(i)
                        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        s.record(stream)
                        e2.eval_ptr(x2.data_ptr(), f2.data_ptr(), g2.data_ptr(), j2.data_ptr(), mode, capi.MEM_DEVICE, sp)
                        e.record(stream)
                        torch.cuda.synchronize()
                        ts.append(s.elapsed_time(e))
                    ms = float(np.median(ts))
                    row[tag] = {"ms_per_batch": ms, "evals_per_s": w2.batch / (ms / 1e3),
                                "frac_of_hbm_peak": unit * w2.batch / (ms / 1e3) / 1e9 / peak}
                others[name] = row
                e2.close()
                del x2, f2, g2, j2
            except Exception as exc:  # noqa: BLE001
                others[name] = {"error": f"{type(exc).__name__}: {exc}"}
        extras["other_configs"] = others

    # ---- C5: all-gather of the full per-instance results [g | Jvals] over NVLink (NCCL), timed on its own ----
    if world > 1 and not args.no_extras:
        try:
            big_g = torch.empty((world * B, ng), dtype=torch.float64, device=dev)
            big_j = torch.empty((world * B, nz), dtype=torch.float64, device=dev)
            for _ in range(2):
                dist.all_gather_into_tensor(big_g, g)
                dist.all_gather_into_tensor(big_j, jac)
            gts = []
            for i in range(3):
                barrier()
                align()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record(stream)
                dist.all_gather_into_tensor(big_g, g)
                dist.all_gather_into_tensor(big_j, jac)
                e.record(stream)
                torch.cuda.synchronize()
                gts.append(s.elapsed_time(e))
            tg = torch.tensor([float(np.mean(gts))], dtype=torch.float64, device=dev)
            dist.all_reduce(tg, op=dist.ReduceOp.MAX)
            gms = float(tg.item())
            recv = 8.0 * (world - 1) * B * (ng + nz)
            ok = bool(torch.equal(big_j[rank * B:(rank + 1) * B], jac) and torch.equal(big_g[rank * B:(rank + 1) * B], g))
            chk = torch.stack([big_j.view(torch.int64).sum(), big_g.view(torch.int64).sum()])
            lo, hi = chk.clone(), chk.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            ok = ok and bool(torch.equal(lo, hi))
            extras["gather_full"] = {
                "ms": gms, "recv_bytes_per_rank": int(recv), "GBps_per_rank": recv / (gms / 1e3) / 1e9,
                "frac_of_nvlink_900GBps": recv / (gms / 1e3) / 1e9 / 900.0,
                "ms_per_step_with_full_gather": total_ms_max / args.steps + gms,
                "evals_per_s_with_full_gather": world * B / ((total_ms_max / args.steps + gms) / 1e3),
                "parity": "bit-equal (own rows == local results; 64-bit checksum of the gathered arrays identical on "
                          "all ranks)" if ok else "MISMATCH",
                "what": f"NCCL all_gather_into_tensor of g[{B}x{ng}] and Jvals[{B}x{nz}] per rank into "
                        f"[{world * B} x ...] on every rank"}
            del big_g, big_j
        except Exception as exc:  # noqa: BLE001
            extras["gather_full"] = {"error": f"{type(exc).__name__}: {exc}"}

    # ---- end to end through the C ABI with HOST buffers (pinned), copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        hx = torch.from_numpy(wl.x).pin_memory()
        hf = torch.empty(B, dtype=torch.float64).pin_memory()
        hg = torch.empty((B, ng), dtype=torch.float64).pin_memory()
        hj = torch.empty((B, nz), dtype=torch.float64).pin_memory()

        def e2e_step():
            ev.eval_ptr(hx.data_ptr(), hf.data_ptr(), hg.data_ptr(), hj.data_ptr(), jac_mode, capi.MEM_HOST, sp)

        nsteps = max(3, min(args.steps, 10))
        for _ in range(2):
            e2e_step()
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(stream)
        for _ in range(nsteps):
            e2e_step()
        e.record(stream)
        barrier()
        te = torch.tensor([s.elapsed_time(e)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": world * B * nsteps / (float(te.item()) / 1e3), "unit": UNIT,
               "h2d_bytes_per_step": int(8 * B * nv), "d2h_bytes_per_step": int(8 * B * (1 + ng + nz)),
               "steps": nsteps, "what": "ecuda_eval with pinned HOST x/f/g/J buffers: H2D of x, kernel, D2H of "
                                        "f, g and all Jacobian values, every step (the library pipelines the call over "
                                        "instance chunks on three streams)"}

    # ---- exact mode end to end: the full triplet array against the compact form (ecuda_eval_compact: only the
    # per-instance triplets cross PCIe; the D-coupled ones are the same for every instance). Same pinned buffers.
    if e2e is not None and not args.no_extras:
        try:
            idx, shared = ev.compact_structure()
            nl = int(idx.size)
            hjl = torch.empty((B, nl), dtype=torch.float64).pin_memory()

            def timed(fn, n):
                for _ in range(2):
                    fn()
                barrier()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record(stream)
                for _ in range(n):
                    fn()
                e.record(stream)
                barrier()
                t = torch.tensor([s.elapsed_time(e)], dtype=torch.float64, device=dev)
                if world > 1:
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                return world * B * n / (float(t.item()) / 1e3)

            n_ec = max(3, min(args.steps, 6))
            v_full = timed(lambda: ev.eval_ptr(hx.data_ptr(), hf.data_ptr(), hg.data_ptr(), hj.data_ptr(), capi.JAC_EXACT,
                                               capi.MEM_HOST, sp), n_ec)
            g_full = hg.numpy()[:64].copy()
            v_comp = timed(lambda: ev.eval_compact_ptr(hx.data_ptr(), hf.data_ptr(), hg.data_ptr(), hjl.data_ptr(),
                                                       capi.MEM_HOST, sp), n_ec)
            spliced = ev.splice(shared, idx, hjl.numpy()[:64])
            same = bool(np.array_equal(spliced, hj.numpy()[:64]) and np.array_equal(hg.numpy()[:64], g_full))
            extras["exact_e2e"] = {
                "full": {"value": v_full, "unit": UNIT, "d2h_bytes_per_step": int(8 * B * (1 + ng + nz))},
                "compact": {"value": v_comp, "unit": UNIT, "d2h_bytes_per_step": int(8 * B * (1 + ng + nl)),
                            "local_triplets": nl, "of": int(nz)},
                "speedup": v_comp / v_full,
                "splice_parity": "bit-equal (first 64 instances spliced on the host)" if same else "MISMATCH",
                "steps": n_ec,
                "what": "ecuda_eval(JAC_EXACT) vs ecuda_eval_compact with the same pinned HOST buffers; the compact "
                        "call runs the same evaluation kernel into device scratch and gathers the per-instance "
                        "triplets on the device (k_gather_local) before the D2H copy"}
            del hjl
        except Exception as exc:  # noqa: BLE001
            extras["exact_e2e"] = {"error": f"{type(exc).__name__}: {exc}"}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline = cpu_sample(wl, args.cpu_seconds, 1 if args.jac == "fd" else 0, style=0)
        tight = cpu_sample(wl, max(2.0, args.cpu_seconds / 4), 1 if args.jac == "fd" else 0, style=1)
        cpu_baseline.pop("seconds"), cpu_baseline.pop("n")
        cpu_baseline["tight_loops_value"] = tight["value"]
        cpu_baseline["note"] = ("oracle port (restatement, not PSOPT/ADOL-C binaries); value = reference-style "
                                "callbacks on all host threads; tight_loops_value = same arithmetic, plain loops; "
                                "tight_rowrestricted_value = plain loops with the row-restricted finite differences "
                                "the GPU kernels use (bit-identical Jacobian, O(nnz) instead of O(groups x ncons) work)")
        if args.jac == "fd" and not args.no_extras:
            try:
                rr = cpu_rowrestricted_sample(wl, max(2.0, args.cpu_seconds / 4))
                cpu_baseline["tight_rowrestricted_value"] = rr["value"]
                cpu_baseline["tight_rowrestricted_sample"] = rr["sample"]
            except Exception as exc:  # noqa: BLE001
                cpu_baseline["tight_rowrestricted_value"] = None
                cpu_baseline["tight_rowrestricted_sample"] = f"unavailable: {type(exc).__name__}: {exc}"

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
                "scaling": "weak" if world == 1 else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(wl, args, world), "clocks": clocks, "e2e": e2e,
                "gpu_launches": int(launches), "roofline": roofline, "exchange": gather_impl}
        if gather_parity:
            line["gather_parity"] = gather_parity
        line.update(extras)
        if e2e and "d2h_probe" in extras:
            # achieved device -> host rate per GPU inside the e2e step against the bare copy measured above
            e2e["frac_of_d2h_probe"] = (e2e["d2h_bytes_per_step"] / (B * world / e2e["value"]) / 1e9) / \
                extras["d2h_probe"]["GBps_per_gpu"]
        if cpu_baseline:
            line["cpu_baseline"] = cpu_baseline
        emit(line)
    ev.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse()
    # exactly one line on stdout (rank 0's JSON): anything a library prints (NCCL's version banner,
    # torchrun notices) goes to stderr while the benchmark runs
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference(args)
    return run_ecuda(args)


_REAL_STDOUT = None


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode())
        sys.stdout.flush()


if __name__ == "__main__":
    sys.exit(main())
