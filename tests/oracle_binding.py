"""ctypes binding of the CPU oracle (oracle/_build/liboracle.so). Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_build", "liboracle.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


class OracleDesc(C.Structure):
    _fields_ = [("model", C.c_int32), ("nphases", C.c_int32), ("nnodes", C.c_int32 * 8),
                ("nstatic", C.c_int32 * 8), ("ncontrols", C.c_int32), ("ntracks", C.c_int32),
                ("nwaypoints", C.c_int32), ("collocation", C.c_int32), ("pattern_mode", C.c_int32),
                ("maximize", C.c_int32), ("batch", C.c_int32), ("index_base", C.c_int32)]


def build_oracle():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)


_lib = None


def lib():
    global _lib
    if _lib is None:
        src_dir = os.path.join(ROOT, "oracle")
        newest = max(os.path.getmtime(os.path.join(src_dir, f)) for f in os.listdir(src_dir)
                     if f.endswith((".cpp", ".hpp")))
        if not os.path.exists(LIB) or (os.path.getmtime(LIB) < newest and os.access(src_dir, os.W_OK)):
            build_oracle()
        L = C.CDLL(LIB)
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.POINTER(OracleDesc)]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_last_error.restype = C.c_char_p
        L.oracle_dims.argtypes = [C.c_void_p, _ip]
        L.oracle_structure.argtypes = [C.c_void_p, _ip, _ip, _ip]
        L.oracle_collocation.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp]
        L.oracle_make_collocation.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp]
        L.oracle_set_scaling.argtypes = [C.c_void_p, _dp, _dp, C.c_double]
        L.oracle_add_border.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, C.c_int]
        L.oracle_add_track.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int, _dp, _dp, _dp]
        L.oracle_set_cylinders.argtypes = [C.c_void_p, _dp]
        L.oracle_edge_geometry.argtypes = [_dp, _dp, _dp]
        L.oracle_eval_batch.restype = C.c_double
        L.oracle_eval_batch.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp, C.c_int,
                                        C.c_int, C.c_int]
        L.oracle_stage_user_rows.argtypes = [C.c_int, _ip]
        L.oracle_stage_user_tape.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _ip, _ip, _ip, _dp, _ip, C.c_int]
        L.oracle_hess_structure.argtypes = [C.c_void_p, _ip, _ip]
        L.oracle_eval_hess.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, _dp, _dp, _dp]
        L.oracle_ode_error.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, _dp]
        L.oracle_resample.argtypes = [C.c_void_p, C.c_int, _dp, _ip, _dp, _dp]
        L.oracle_max_threads.restype = C.c_int
        L.oracle_sincos.argtypes = [_dp, C.c_int, _dp, _dp]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def make_collocation(kind, N):
    tau, w, D = np.zeros(N), np.zeros(N), np.zeros((N, N))
    if lib().oracle_make_collocation(kind, N, _p(tau), _p(w), _p(D)) != 0:
        raise RuntimeError(lib().oracle_last_error().decode())
    return tau, w, D


def edge_geometry(a, b):
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    out = np.zeros(6)
    lib().oracle_edge_geometry(_p(a), _p(b), _p(out))
    return out


class Oracle:
    """One oracle problem + a batch of raw VGP instances (see etol_b200.workloads.Workload)."""

    def __init__(self, workload):
        wl = workload
        d = OracleDesc()
        d.model, d.nphases = wl.model, wl.nphases
        if getattr(wl, "tape", None) is not None:  # user model: the oracle replays the same tape (model 3)
            t = wl.tape
            ops = np.array([n[0] for n in t.nodes], dtype=np.int32)
            aa = np.array([n[1] for n in t.nodes], dtype=np.int32)
            bb = np.array([n[2] for n in t.nodes], dtype=np.int32)
            imm = np.array([n[3] for n in t.nodes], dtype=np.float64)
            fo = np.array(t.f_out, dtype=np.int32)
            lib().oracle_stage_user_tape(t.ns, t.nc, int(t.static_kind == 1), len(t.nodes), ops.ctypes.data_as(_ip),
                                         aa.ctypes.data_as(_ip), bb.ctypes.data_as(_ip), _p(imm),
                                         fo.ctypes.data_as(_ip), t.cost_out)
            ro = np.array(getattr(t, "row_out", []), dtype=np.int32)
            lib().oracle_stage_user_rows(len(ro), ro.ctypes.data_as(_ip))
            d.model = 3
        for p in range(wl.nphases):
            d.nnodes[p] = wl.nnodes[p]
            d.nstatic[p] = wl.nstatic[p]
        d.ncontrols, d.ntracks, d.nwaypoints = wl.ncontrols, wl.ntracks, wl.nwaypoints
        d.collocation, d.pattern_mode = wl.collocation, wl.pattern_mode
        d.maximize, d.batch, d.index_base = int(wl.maximize), wl.batch, wl.index_base
        self.h = lib().oracle_create(C.byref(d))
        if not self.h:
            raise RuntimeError(lib().oracle_last_error().decode())
        dims = np.zeros(7, dtype=np.int32)
        lib().oracle_dims(self.h, dims.ctypes.data_as(_ip))
        (self.nvars, self.ncons, self.nnz, self.ngroups, self.ns, self.nc, self.nlink) = [int(v) for v in dims]
        self.batch = wl.batch
        # raw data
        if wl.borders is not None:
            for b in range(wl.batch):
                for p in range(wl.nphases):
                    for poly in wl.borders[b][p]:
                        arr = np.ascontiguousarray(poly, dtype=np.float64)
                        lib().oracle_add_border(self.h, b, p, _p(arr), arr.shape[0])
        if wl.tracks is not None:
            for b in range(wl.batch):
                for (radius, t, x, y) in wl.tracks[b]:
                    t = np.ascontiguousarray(t, dtype=np.float64)
                    x = np.ascontiguousarray(x, dtype=np.float64)
                    y = np.ascontiguousarray(y, dtype=np.float64)
                    lib().oracle_add_track(self.h, b, float(radius), len(t), _p(t), _p(x), _p(y))
        if wl.cylinders is not None:
            cyl = np.ascontiguousarray(wl.cylinders, dtype=np.float64)
            lib().oracle_set_cylinders(self.h, _p(cyl))
        if wl.sz is not None or wl.sg is not None or wl.sf != 1.0:
            self.set_scaling(wl.sz, wl.sg, wl.sf)

    def __del__(self):
        if getattr(self, "h", None):
            lib().oracle_destroy(self.h)
            self.h = None

    def set_scaling(self, sz, sg, sf):
        sz = None if sz is None else np.ascontiguousarray(sz, dtype=np.float64)
        sg = None if sg is None else np.ascontiguousarray(sg, dtype=np.float64)
        lib().oracle_set_scaling(self.h, _p(sz), _p(sg), float(sf))

    def structure(self):
        irow = np.zeros(self.nnz, dtype=np.int32)
        jcol = np.zeros(self.nnz, dtype=np.int32)
        grp = np.zeros(self.nvars, dtype=np.int32)
        lib().oracle_structure(self.h, irow.ctypes.data_as(_ip), jcol.ctypes.data_as(_ip),
                               grp.ctypes.data_as(_ip))
        return irow, jcol, grp

    def collocation(self, phase, N):
        tau, w, D = np.zeros(N), np.zeros(N), np.zeros((N, N))
        lib().oracle_collocation(self.h, phase, _p(tau), _p(w), _p(D))
        return tau, w, D

    def eval(self, x, want=("f", "g", "jac"), jac_mode=1, style=0, nthreads=1, first=0, count=None):
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, self.nvars)
        count = x.shape[0] if count is None else count
        out = {}
        f = np.zeros(count) if "f" in want else None
        g = np.zeros((count, self.ncons)) if "g" in want else None
        jac = np.zeros((count, self.nnz)) if "jac" in want else None
        grad = np.zeros((count, self.nvars)) if "grad" in want else None
        secs = lib().oracle_eval_batch(self.h, first, count, _p(x), _p(f), _p(g), _p(jac), _p(grad),
                                       jac_mode, style, nthreads)
        if secs < 0:
            raise RuntimeError(lib().oracle_last_error().decode())
        out.update(f=f, g=g, jac=jac, grad=grad, seconds=secs)
        return out


def hess_structure(o):
    n = lib().oracle_hess_structure(o.h, None, None)
    irow, jcol = np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
    lib().oracle_hess_structure(o.h, irow.ctypes.data_as(_ip), jcol.ctypes.data_as(_ip))
    return irow, jcol


def eval_hess(o, x, sigma, lam):
    """Hessian of the Lagrangian sigma*f + lam.g (scaled space), lower triangle, [B][nnz_h]"""
    x = np.ascontiguousarray(x, dtype=np.float64)
    sigma = np.ascontiguousarray(sigma, dtype=np.float64)
    lam = np.ascontiguousarray(lam, dtype=np.float64)
    out = np.zeros((x.shape[0], lib().oracle_hess_structure(o.h, None, None)))
    if lib().oracle_eval_hess(o.h, 0, x.shape[0], _p(x), _p(sigma), _p(lam), _p(out)) != 0:
        raise RuntimeError(lib().oracle_last_error().decode())
    return out


def ode_error(o, wl, x):
    """relative local discretisation error per mesh interval, [B][sum_p (N_p - 1)]"""
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.zeros((x.shape[0], sum(n - 1 for n in wl.nnodes)))
    if lib().oracle_ode_error(o.h, 0, x.shape[0], _p(x), _p(out)) != 0:
        raise RuntimeError(lib().oracle_last_error().decode())
    return out


def resample(o, wl, x, nnodes_new, sz_new=None):
    x = np.ascontiguousarray(x, dtype=np.float64)
    nn = np.asarray(nnodes_new, dtype=np.int32)
    nv = sum((wl.ns + wl.nc) * int(n) + 2 for n in nn)
    out = np.zeros((x.shape[0], nv))
    sz = None if sz_new is None else np.ascontiguousarray(sz_new, dtype=np.float64)
    if lib().oracle_resample(o.h, x.shape[0], _p(x), nn.ctypes.data_as(_ip), _p(sz), _p(out)) != 0:
        raise RuntimeError(lib().oracle_last_error().decode())
    return out


def max_threads():
    return lib().oracle_max_threads()


def det_sincos(x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    s, c = np.zeros_like(x), np.zeros_like(x)
    lib().oracle_sincos(_p(x), x.size, _p(s), _p(c))
    return s, c
