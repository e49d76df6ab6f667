"""ctypes binding of the eCUDA C ABI (include/ecuda.h -> etol_b200/csrc/libecuda.so).

This is the Python mirror of the boundary a C++ host (src/eCUDA) links against; tests and bench.py
drive the CUDA path through it. There is no fallback of any kind: if the shared library is missing
the import of `lib()` raises, and if no sm_100 GPU is usable `Evaluator(...)` raises EcudaError.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# ECUDA_LIB: alternative build of the same library (kernel tuning experiments); still no fallback
LIB_PATH = os.environ.get("ECUDA_LIB") or os.path.join(HERE, "csrc", "libecuda.so")

MAX_PHASES = 8
MEM_HOST, MEM_DEVICE = 0, 1
JAC_EXACT, JAC_FD = 0, 1

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


class ProblemDesc(C.Structure):
    _fields_ = [("model", C.c_int32), ("nphases", C.c_int32), ("nnodes", C.c_int32 * MAX_PHASES),
                ("nstatic", C.c_int32 * MAX_PHASES), ("ncontrols", C.c_int32), ("ntracks", C.c_int32),
                ("nwaypoints", C.c_int32), ("collocation", C.c_int32), ("pattern_mode", C.c_int32),
                ("maximize", C.c_int32), ("batch", C.c_int32), ("index_base", C.c_int32)]


class Dims(C.Structure):
    _fields_ = [("nvars", C.c_int32), ("ncons", C.c_int32), ("nnz", C.c_int32), ("ngroups", C.c_int32),
                ("nstates", C.c_int32), ("ncontrols", C.c_int32), ("nlinkages", C.c_int32),
                ("inst_stride", C.c_int32), ("rec_size", C.c_int32), ("track_size", C.c_int32)]


class TapeNode(C.Structure):
    _fields_ = [("op", C.c_int32), ("a", C.c_int32), ("b", C.c_int32), ("reserved", C.c_int32), ("imm", C.c_double)]


class UserModel(C.Structure):
    _fields_ = [("nstates", C.c_int32), ("ncontrols", C.c_int32), ("static_kind", C.c_int32), ("nnodes", C.c_int32),
                ("nodes", C.POINTER(TapeNode)), ("f_out", C.c_int32 * 8), ("cost_out", C.c_int32)]


class EcudaError(RuntimeError):
    pass


# every symbol include/ecuda.h declares (tests/test_abi.py checks the library exports them all)
ABI_SYMBOLS = [
    "ecuda_abi_version", "ecuda_create", "ecuda_destroy", "ecuda_last_error", "ecuda_set_problem",
    "ecuda_get_dims", "ecuda_get_structure", "ecuda_get_collocation", "ecuda_set_collocation",
    "ecuda_set_scaling", "ecuda_upload_instances", "ecuda_upload_bounds", "ecuda_eval", "ecuda_eval_grad_f",
    "ecuda_get_compact_structure", "ecuda_eval_compact", "ecuda_splice_jacobian",
    "ecuda_summary", "ecuda_summarize", "ecuda_summarize_allgather", "ecuda_eval_allgather", "ecuda_peer_barrier", "ecuda_peer_barrier_status", "ecuda_sync", "ecuda_launch_count", "ecuda_fp64_peak", "ecuda_ipopt_eval_f", "ecuda_ipopt_eval_grad_f",
    "ecuda_ipopt_eval_g", "ecuda_ipopt_eval_jac_g", "ecuda_set_ipopt_jac_mode", "ecuda_si2d_edge_records",
    "ecuda_host_dims", "ecuda_host_structure", "ecuda_host_compact_structure", "ecuda_host_collocation", "ecuda_host_model_eval",
    "ecuda_host_path_eval", "ecuda_get_hess_structure", "ecuda_eval_hess", "ecuda_ipopt_eval_h", "ecuda_host_hess_structure",
    "ecuda_ode_error", "ecuda_resample", "ecuda_host_error_mesh", "ecuda_host_resample_matrix",
    "ecuda_register_user_model", "ecuda_register_user_model_rows", "ecuda_user_model_source", "ecuda_user_model_compile_check",
]

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EcudaError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                         f"g.build()'` (nvcc, sm_100a). eCUDA has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    L.ecuda_abi_version.restype = C.c_int
    L.ecuda_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    L.ecuda_destroy.argtypes = [C.c_void_p]
    L.ecuda_last_error.restype = C.c_char_p
    L.ecuda_last_error.argtypes = [C.c_void_p]
    L.ecuda_set_problem.argtypes = [C.c_void_p, C.POINTER(ProblemDesc)]
    L.ecuda_get_dims.argtypes = [C.c_void_p, C.POINTER(Dims)]
    L.ecuda_get_structure.argtypes = [C.c_void_p, _ip, _ip, _ip]
    L.ecuda_get_collocation.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp]
    L.ecuda_set_collocation.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp]
    L.ecuda_set_scaling.argtypes = [C.c_void_p, _dp, _dp, C.c_double]
    L.ecuda_upload_instances.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.ecuda_upload_bounds.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    L.ecuda_eval.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                             C.c_void_p]
    L.ecuda_get_compact_structure.argtypes = [C.c_void_p, _ip, _ip, _dp]
    L.ecuda_eval_compact.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.ecuda_splice_jacobian.argtypes = [_dp, _ip, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]
    L.ecuda_eval_grad_f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.ecuda_summary.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.ecuda_summarize.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.ecuda_summarize_allgather.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_int,
                                            C.c_void_p]
    L.ecuda_eval_allgather.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                       C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_void_p]
    L.ecuda_peer_barrier.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_uint64, C.c_void_p]
    L.ecuda_peer_barrier_status.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_uint64),
                                            C.POINTER(C.c_int32), C.c_int]
    L.ecuda_sync.argtypes = [C.c_void_p]
    L.ecuda_launch_count.restype = C.c_int64
    L.ecuda_launch_count.argtypes = [C.c_void_p]
    L.ecuda_fp64_peak.argtypes = [C.c_void_p, _dp]
    L.ecuda_ipopt_eval_f.argtypes = [C.c_void_p, C.c_int, _dp, C.c_int, _dp]
    L.ecuda_ipopt_eval_grad_f.argtypes = [C.c_void_p, C.c_int, _dp, C.c_int, _dp]
    L.ecuda_ipopt_eval_g.argtypes = [C.c_void_p, C.c_int, _dp, C.c_int, C.c_int, _dp]
    L.ecuda_ipopt_eval_jac_g.argtypes = [C.c_void_p, C.c_int, _dp, C.c_int, C.c_int, C.c_int, _ip, _ip, _dp]
    L.ecuda_set_ipopt_jac_mode.argtypes = [C.c_void_p, C.c_int]
    L.ecuda_si2d_edge_records.argtypes = [_dp, C.c_int, _dp]
    L.ecuda_host_dims.argtypes = [C.POINTER(ProblemDesc), C.POINTER(Dims)]
    L.ecuda_host_structure.argtypes = [C.POINTER(ProblemDesc), _ip, _ip, _ip]
    L.ecuda_host_compact_structure.argtypes = [C.POINTER(ProblemDesc), _ip, _ip]
    L.ecuda_host_collocation.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp]
    L.ecuda_host_model_eval.argtypes = [C.c_int, _dp, _dp, C.c_double, _dp, _dp]
    L.ecuda_get_hess_structure.argtypes = [C.c_void_p, _ip, _ip, _ip]
    L.ecuda_host_hess_structure.argtypes = [C.POINTER(ProblemDesc), _ip, _ip, _ip]
    L.ecuda_eval_hess.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.ecuda_ipopt_eval_h.argtypes = [C.c_void_p, C.c_int, _dp, C.c_int, C.c_double, C.c_int, _dp, C.c_int, C.c_int, _ip, _ip, _dp]
    L.ecuda_ode_error.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.ecuda_resample.argtypes = [C.c_void_p, C.c_void_p, _ip, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.ecuda_host_error_mesh.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, _dp]
    L.ecuda_host_resample_matrix.argtypes = [C.c_int, C.c_int, C.c_int, _dp]
    L.ecuda_register_user_model_rows.argtypes = [C.POINTER(UserModel), C.c_int32, _ip, _ip, C.c_char_p, C.c_size_t]
    L.ecuda_register_user_model.argtypes = [C.POINTER(UserModel), _ip, C.c_char_p, C.c_size_t]
    L.ecuda_user_model_source.argtypes = [C.c_int32, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.ecuda_user_model_compile_check.argtypes = [C.c_int32, C.c_int, C.POINTER(C.c_size_t), C.c_char_p, C.c_size_t]
    _lib = L
    return L


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


_registered = {}


def register_user_model(tape):
    """etol_b200.tape.Tape -> model id for ProblemDesc.model (one registration per distinct tape and process)."""
    key = tape.key()
    if key in _registered:
        return _registered[key]
    nodes = (TapeNode * len(tape.nodes))()
    for i, (op, a, b, imm) in enumerate(tape.nodes):
        nodes[i].op, nodes[i].a, nodes[i].b, nodes[i].imm = op, a, b, imm
    um = UserModel()
    um.nstates, um.ncontrols, um.static_kind, um.nnodes = tape.ns, tape.nc, tape.static_kind, len(tape.nodes)
    um.nodes = nodes
    for i, v in enumerate(tape.f_out):
        um.f_out[i] = v
    um.cost_out = tape.cost_out
    mid = C.c_int32(-1)
    err = C.create_string_buffer(512)
    rows = np.array(getattr(tape, "row_out", []), dtype=np.int32)
    if rows.size:  # traced path rows
        rc = lib().ecuda_register_user_model_rows(C.byref(um), int(rows.size), rows.ctypes.data_as(_ip), C.byref(mid), err,
                                                  len(err))
    else:
        rc = lib().ecuda_register_user_model(C.byref(um), C.byref(mid), err, len(err))
    if rc != 0:
        raise EcudaError("ecuda_register_user_model: " + err.value.decode())
    _registered[key] = mid.value
    return mid.value


def user_model_source(model_id):
    n = C.c_size_t(0)
    if lib().ecuda_user_model_source(model_id, None, 0, C.byref(n)) != 0:
        raise EcudaError("not a registered user model")
    buf = C.create_string_buffer(n.value)
    lib().ecuda_user_model_source(model_id, buf, n.value, None)
    return buf.value.decode()


def user_model_compile_check(model_id, nnodes):
    """NVRTC build of the kernels for a registered model (no device needed) -> size of the sm_100a image."""
    n = C.c_size_t(0)
    log = C.create_string_buffer(1 << 16)
    rc = lib().ecuda_user_model_compile_check(model_id, nnodes, C.byref(n), log, len(log))
    if rc != 0:
        raise EcudaError("user model did not compile: " + log.value.decode())
    return n.value


def host_model_eval(model, x, u, t=0.0):
    x = np.ascontiguousarray(x, dtype=np.float64)
    u = np.ascontiguousarray(u, dtype=np.float64)
    f, cost = np.zeros(8), np.zeros(1)
    if lib().ecuda_host_model_eval(model, _p(x), _p(u), float(t), _p(f), _p(cost)) != 0:
        raise EcudaError("ecuda_host_model_eval: unknown model")
    return f[:len(x)].copy(), float(cost[0])


def make_desc(wl):
    d = ProblemDesc()
    d.model, d.nphases = wl.model, wl.nphases
    for p in range(wl.nphases):
        d.nnodes[p] = wl.nnodes[p]
        d.nstatic[p] = wl.nstatic[p]
    d.ncontrols, d.ntracks, d.nwaypoints = wl.ncontrols, wl.ntracks, wl.nwaypoints
    d.collocation, d.pattern_mode = wl.collocation, wl.pattern_mode
    d.maximize, d.batch, d.index_base = int(wl.maximize), wl.batch, wl.index_base
    return d


def host_dims(wl):
    d, out = make_desc(wl), Dims()
    if lib().ecuda_host_dims(C.byref(d), C.byref(out)) != 0:
        raise EcudaError("ecuda_host_dims rejected the problem description")
    return out


def host_structure(wl):
    d, dm = make_desc(wl), host_dims(wl)
    irow = np.zeros(dm.nnz, dtype=np.int32)
    jcol = np.zeros(dm.nnz, dtype=np.int32)
    grp = np.zeros(dm.nvars, dtype=np.int32)
    rc = lib().ecuda_host_structure(C.byref(d), irow.ctypes.data_as(_ip), jcol.ctypes.data_as(_ip),
                                    grp.ctypes.data_as(_ip))
    if rc != 0:
        raise EcudaError("ecuda_host_structure failed")
    return irow, jcol, grp


def host_compact_structure(wl):
    """ascending triplet indices of the per-instance part of an exact Jacobian (ecuda_eval_compact)"""
    d = make_desc(wl)
    n = C.c_int32(0)
    if lib().ecuda_host_compact_structure(C.byref(d), C.byref(n), None):
        raise EcudaError("ecuda_host_compact_structure failed")
    idx = np.zeros(n.value, dtype=np.int32)
    lib().ecuda_host_compact_structure(C.byref(d), None, idx.ctypes.data_as(_ip))
    return idx


def splice_jacobian(shared, idx, jac_local, nnz):
    jac_local = np.ascontiguousarray(jac_local, dtype=np.float64)
    shared = np.ascontiguousarray(shared, dtype=np.float64)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    full = np.empty((jac_local.shape[0], nnz))
    rc = lib().ecuda_splice_jacobian(_p(shared), idx.ctypes.data_as(_ip), nnz, len(idx), jac_local.ctypes.data,
                                     jac_local.shape[0], full.ctypes.data)
    if rc:
        raise EcudaError(f"ecuda_splice_jacobian: {rc}")
    return full


def host_hess_structure(wl):
    """lower triangle of the Lagrangian Hessian: (iRow, jCol), sorted by (column, row)"""
    d, n = make_desc(wl), C.c_int32(0)
    if lib().ecuda_host_hess_structure(C.byref(d), C.byref(n), None, None) != 0:
        raise EcudaError("ecuda_host_hess_structure rejected the problem description")
    irow, jcol = np.zeros(n.value, dtype=np.int32), np.zeros(n.value, dtype=np.int32)
    lib().ecuda_host_hess_structure(C.byref(d), None, irow.ctypes.data_as(_ip), jcol.ctypes.data_as(_ip))
    return irow, jcol


def host_collocation(kind, N):
    tau, w, D = np.zeros(N), np.zeros(N), np.zeros((N, N))
    if lib().ecuda_host_collocation(kind, N, _p(tau), _p(w), _p(D)) != 0:
        raise EcudaError("ecuda_host_collocation failed")
    return tau, w, D


def edge_records(corners_xy):
    """Polygon corners (n,2) -> (n,6) edge records, via the library's host helper."""
    c = np.ascontiguousarray(corners_xy, dtype=np.float64)
    out = np.zeros((c.shape[0], 6))
    if lib().ecuda_si2d_edge_records(_p(c), c.shape[0], _p(out)) != 0:
        raise EcudaError("ecuda_si2d_edge_records failed")
    return out


def host_path_eval(wl, inst_row, x, y, t):
    """path rows of phase 0 of one instance block at position (x, y), time t (static, moving-zone, traced rows)"""
    d = make_desc(wl)
    inst_row = np.ascontiguousarray(inst_row, dtype=np.float64)
    rows = np.zeros(wl.npath[0])
    L = lib()
    L.ecuda_host_path_eval.argtypes = [C.POINTER(ProblemDesc), _dp, C.c_double, C.c_double, C.c_double, _dp]
    if L.ecuda_host_path_eval(C.byref(d), _p(inst_row), float(x), float(y), float(t), _p(rows)) != 0:
        raise EcudaError("ecuda_host_path_eval failed")
    return rows


def pack_instances(wl, dims=None):
    """Raw VGP data of a Workload -> the [B][inst_stride] instance-data block of ecuda_upload_instances."""
    dims = dims or host_dims(wl)
    B = wl.batch
    out = np.zeros((B, dims.inst_stride))
    off = 0
    if wl.borders is not None:  # polygons -> edge records (host helper of the library), then tracks
        for b in range(B):
            o = 0
            for p in range(wl.nphases):
                for poly in wl.borders[b][p]:
                    rec = edge_records(np.asarray(poly)[:, :2])
                    out[b, o:o + rec.size] = rec.ravel()
                    o += rec.size
            for (radius, t, x, y) in (wl.tracks[b] if wl.tracks else []):
                rec = np.concatenate([[radius], np.stack([t, x, y], axis=1).ravel()])
                out[b, o:o + rec.size] = rec
                o += rec.size
    else:  # cylinders: cx, cy, r*r, 0
        cyl = np.asarray(wl.cylinders)
        n = cyl.shape[1]
        rec = np.zeros((B, n, 4))
        rec[:, :, 0] = cyl[:, :, 0]
        rec[:, :, 1] = cyl[:, :, 1]
        rec[:, :, 2] = cyl[:, :, 2] * cyl[:, :, 2]
        out[:, off:off + 4 * n] = rec.reshape(B, 4 * n)
        for b in range(B):
            o = off + 4 * n
            for (radius, t, x, y) in (wl.tracks[b] if wl.tracks else []):
                trk = np.concatenate([[radius], np.stack([t, x, y], axis=1).ravel()])
                out[b, o:o + trk.size] = trk
                o += trk.size
    return np.ascontiguousarray(out)


class Evaluator:
    """One ecuda handle: problem structure + a batch of instances resident on one GPU."""

    def __init__(self, wl, device=0, upload=True):
        self.L = lib()
        self.h = C.c_void_p()
        rc = self.L.ecuda_create(device, C.byref(self.h))
        if rc != 0:
            msg = self.L.ecuda_last_error(None).decode()
            self.h = None
            raise EcudaError(f"ecuda_create failed ({rc}): {msg}")
        self.wl = wl
        desc = make_desc(wl)
        self._check(self.L.ecuda_set_problem(self.h, C.byref(desc)))
        self.dims = Dims()
        self._check(self.L.ecuda_get_dims(self.h, C.byref(self.dims)))
        self.nvars, self.ncons, self.nnz = self.dims.nvars, self.dims.ncons, self.dims.nnz
        self.batch = wl.batch
        if wl.sz is not None or wl.sg is not None or wl.sf != 1.0:
            self.set_scaling(wl.sz, wl.sg, wl.sf)
        if upload:
            inst = pack_instances(wl, self.dims)
            self._check(self.L.ecuda_upload_instances(self.h, inst.ctypes.data, MEM_HOST))
            if wl.gl is not None and wl.gu is not None:
                gl = np.ascontiguousarray(wl.gl, dtype=np.float64)
                gu = np.ascontiguousarray(wl.gu, dtype=np.float64)
                # +-inf bounds are kept as they are: violation arithmetic handles them
                self._check(self.L.ecuda_upload_bounds(self.h, gl.ctypes.data, gu.ctypes.data, MEM_HOST))

    def _check(self, rc):
        if rc != 0:
            raise EcudaError(f"ecuda error {rc}: {self.L.ecuda_last_error(self.h).decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.L.ecuda_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_scaling(self, sz, sg, sf=1.0):
        sz = None if sz is None else np.ascontiguousarray(sz, dtype=np.float64)
        sg = None if sg is None else np.ascontiguousarray(sg, dtype=np.float64)
        self._check(self.L.ecuda_set_scaling(self.h, _p(sz), _p(sg), float(sf)))

    def structure(self):
        irow = np.zeros(self.nnz, dtype=np.int32)
        jcol = np.zeros(self.nnz, dtype=np.int32)
        grp = np.zeros(self.nvars, dtype=np.int32)
        self._check(self.L.ecuda_get_structure(self.h, irow.ctypes.data_as(_ip), jcol.ctypes.data_as(_ip),
                                               grp.ctypes.data_as(_ip)))
        return irow, jcol, grp

    def collocation(self, phase):
        N = self.wl.nnodes[phase]
        tau, w, D = np.zeros(N), np.zeros(N), np.zeros((N, N))
        self._check(self.L.ecuda_get_collocation(self.h, phase, _p(tau), _p(w), _p(D)))
        return tau, w, D

    def set_collocation(self, phase, tau, w, D):
        tau, w, D = (np.ascontiguousarray(a, dtype=np.float64) for a in (tau, w, D))
        self._check(self.L.ecuda_set_collocation(self.h, phase, _p(tau), _p(w), _p(D)))

    # ---- host-buffer evaluation (numpy in / numpy out; copies inside the call) ----------------------
    def eval_host(self, x, want=("f", "g", "jac"), jac_mode=JAC_FD):
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(self.batch, self.nvars)
        f = np.zeros(self.batch) if "f" in want else None
        g = np.zeros((self.batch, self.ncons)) if "g" in want else None
        jac = np.zeros((self.batch, self.nnz)) if "jac" in want else None
        ptr = lambda a: None if a is None else a.ctypes.data
        self._check(self.L.ecuda_eval(self.h, x.ctypes.data, ptr(f), ptr(g), ptr(jac), jac_mode, MEM_HOST, None))
        out = dict(f=f, g=g, jac=jac, grad=None)
        if "grad" in want:
            grad = np.zeros((self.batch, self.nvars))
            self._check(self.L.ecuda_eval_grad_f(self.h, x.ctypes.data, grad.ctypes.data, MEM_HOST, None))
            out["grad"] = grad
        return out

    # ---- compact exact Jacobian: per-instance triplets only (ecuda.h "compact exact Jacobian") ---------
    def compact_structure(self):
        """(local_index [nlocal] int32, shared_vals [nnz])"""
        n = C.c_int32(0)
        self._check(self.L.ecuda_get_compact_structure(self.h, C.byref(n), None, None))
        idx = np.zeros(n.value, dtype=np.int32)
        shared = np.zeros(self.nnz)
        self._check(self.L.ecuda_get_compact_structure(self.h, None, idx.ctypes.data_as(_ip), _p(shared)))
        return idx, shared

    def eval_compact_host(self, x, nlocal):
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(self.batch, self.nvars)
        f, g, jl = np.zeros(self.batch), np.zeros((self.batch, self.ncons)), np.zeros((self.batch, nlocal))
        self._check(self.L.ecuda_eval_compact(self.h, x.ctypes.data, f.ctypes.data, g.ctypes.data, jl.ctypes.data, MEM_HOST, None))
        return dict(f=f, g=g, jac_local=jl)

    def eval_compact_ptr(self, x_ptr, f_ptr, g_ptr, jl_ptr, memkind, stream=None):
        self._check(self.L.ecuda_eval_compact(self.h, x_ptr, f_ptr, g_ptr, jl_ptr, memkind, stream))

    def splice(self, shared, idx, jac_local):
        return splice_jacobian(shared, idx, jac_local, self.nnz)

    # ---- raw-pointer evaluation (device tensors or pinned host tensors from torch) -----------------
    def eval_ptr(self, x_ptr, f_ptr, g_ptr, jac_ptr, jac_mode, memkind, stream=None):
        self._check(self.L.ecuda_eval(self.h, x_ptr, f_ptr, g_ptr, jac_ptr, jac_mode, memkind, stream))

    def summary_ptr(self, x_ptr, out_ptr, memkind, stream=None):
        self._check(self.L.ecuda_summary(self.h, x_ptr, out_ptr, memkind, stream))

    def summarize_ptr(self, f_ptr, g_ptr, out_ptr, stream=None):
        self._check(self.L.ecuda_summarize(self.h, f_ptr, g_ptr, out_ptr, stream))

    def summarize_allgather_ptr(self, f_ptr, g_ptr, peer_ptrs, rank, stream=None):
        arr = (C.c_void_p * len(peer_ptrs))(*peer_ptrs)
        self._check(self.L.ecuda_summarize_allgather(self.h, f_ptr, g_ptr, arr, len(peer_ptrs), rank, stream))

    def eval_allgather_ptr(self, x_ptr, f_ptr, g_ptr, jac_ptr, jac_mode, peer_ptrs, rank, stream=None):
        arr = (C.c_void_p * len(peer_ptrs))(*peer_ptrs)
        self._check(self.L.ecuda_eval_allgather(self.h, x_ptr, f_ptr, g_ptr, jac_ptr, jac_mode, arr, len(peer_ptrs), rank,
                                                stream))

    def peer_barrier_ptr(self, flag_ptrs, rank, step, stream=None):
        arr = (C.c_void_p * len(flag_ptrs))(*flag_ptrs)
        self._check(self.L.ecuda_peer_barrier(self.h, arr, len(flag_ptrs), rank, step, stream))

    def peer_barrier_status(self, stream=None, reset=False):
        """(timed_out, step, late_rank) of the peer barriers issued through this handle (sticky until reset)"""
        to, st, lr = C.c_int32(0), C.c_uint64(0), C.c_int32(-1)
        self._check(self.L.ecuda_peer_barrier_status(self.h, stream, C.byref(to), C.byref(st), C.byref(lr), int(reset)))
        return bool(to.value), int(st.value), int(lr.value)

    def sync_status(self):
        """ecuda_sync return code (0, or ECUDA_ERR_PEER after a peer barrier timed out) and the error text"""
        rc = self.L.ecuda_sync(self.h)
        return rc, (self.L.ecuda_last_error(self.h) or b"").decode()

    def hess_host(self, x, sigma, lam):
        """Hessian of the Lagrangian per instance (host buffers): sigma [B], lam [B][ncons] -> [B][nnz_h]"""
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(self.batch, self.nvars)
        sigma = np.ascontiguousarray(sigma, dtype=np.float64).reshape(self.batch)
        lam = np.ascontiguousarray(lam, dtype=np.float64).reshape(self.batch, self.ncons)
        n = C.c_int32(0)
        self._check(lib().ecuda_get_hess_structure(self.h, C.byref(n), None, None))
        out = np.zeros((self.batch, n.value))
        self._check(lib().ecuda_eval_hess(self.h, x.ctypes.data, sigma.ctypes.data, 0.0, lam.ctypes.data, out.ctypes.data,
                                          MEM_HOST, None))
        return out

    def ode_error_host(self, x):
        """relative local discretisation error per mesh interval, [B][sum_p (N_p - 1)] (host buffers)"""
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(self.batch, self.nvars)
        out = np.zeros((self.batch, sum(n - 1 for n in self.wl.nnodes)))
        self._check(lib().ecuda_ode_error(self.h, x.ctypes.data, out.ctypes.data, MEM_HOST, None))
        return out

    def resample_host(self, x, nnodes_new, sz_new=None):
        """decision vectors interpolated onto meshes of nnodes_new[p] nodes (host buffers)"""
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(self.batch, self.nvars)
        nn = np.asarray(nnodes_new, dtype=np.int32)
        nv = sum((self.wl.ns + self.wl.nc) * int(n) + 2 for n in nn)
        out = np.zeros((self.batch, nv))
        sz = None if sz_new is None else np.ascontiguousarray(sz_new, dtype=np.float64)
        self._check(lib().ecuda_resample(self.h, x.ctypes.data, nn.ctypes.data_as(_ip), None if sz is None else sz.ctypes.data,
                                         out.ctypes.data, MEM_HOST, None))
        return out

    def summary_host(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(self.batch, self.nvars)
        out = np.zeros((self.batch, 2))
        self._check(self.L.ecuda_summary(self.h, x.ctypes.data, out.ctypes.data, MEM_HOST, None))
        return out

    def sync(self):
        self._check(self.L.ecuda_sync(self.h))

    def launch_count(self):
        return int(self.L.ecuda_launch_count(self.h))

    def fp64_peak_tflops(self):
        out = C.c_double()
        self._check(self.L.ecuda_fp64_peak(self.h, C.byref(out)))
        return out.value
