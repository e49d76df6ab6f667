"""ctypes binding of build/libetol_test_shim.so (tests/plugin/shim.cpp): the C++ plugin layer
(core-lite TrajectoryOptimizer, eCUDA, NLP drivers) as seen from the python test-suite."""
import ctypes as C
import os
import subprocess

import numpy as np

from etol_b200 import capi, workloads as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "build", "libetol_test_shim.so")
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
EVAL_CB = C.CFUNCTYPE(C.c_int, _dp, _dp, _dp, _dp, _dp)
_lib = None


def lib():
    global _lib
    if _lib is None:
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "src")], check=True)
        L = C.CDLL(SHIM)
        L.shim_create.restype = C.c_void_p
        L.shim_destroy.argtypes = [C.c_void_p]
        L.shim_load.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_char_p]
        L.shim_load_callbacks.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_char_p]
        L.shim_vgp.argtypes = [C.c_void_p, C.POINTER(C.c_int), _dp]
        L.shim_dims.argtypes = [C.c_void_p, C.POINTER(capi.Dims)]
        L.shim_desc.argtypes = [C.c_void_p, C.POINTER(capi.ProblemDesc)]
        L.shim_bounds.argtypes = [C.c_void_p] + [_dp] * 7
        L.shim_instance.argtypes = [C.c_void_p, C.c_int, _dp]
        L.shim_structure.argtypes = [C.c_void_p, _ip, _ip, _ip]
        L.shim_gen_region.argtypes = [_dp, C.c_int, _dp, C.c_int]
        L.shim_num_partitioned_zones.argtypes = [C.c_void_p]
        L.shim_edit_instance_and_remesh.restype = C.c_double
        L.shim_edit_instance_and_remesh.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int]
        L.shim_save_xml.argtypes = [C.c_void_p, C.c_char_p]
        L.shim_save_csv.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_char_p, C.c_int]
        L.shim_interp.restype = C.c_double
        L.shim_interp.argtypes = [C.c_double, C.c_int, _dp, _dp]
        L.shim_nlp_solve.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _ip, _ip, EVAL_CB, C.c_int,
                                     C.c_double, C.c_int, _dp, _dp]
        L.shim_setup.argtypes = [C.c_void_p]
        L.shim_evaluate.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp]
        L.shim_solve.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, C.POINTER(C.c_int), _dp]
        L.shim_traj.argtypes = [C.c_void_p, C.c_int, _dp]
        L.shim_set_hessian.argtypes = [C.c_void_p, C.c_char_p]
        L.shim_set_mesh.argtypes = [C.c_void_p, C.c_char_p, C.c_double, C.c_int]
        L.shim_mesh_history.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), _dp]
        L.shim_next_mesh_size.argtypes = [C.c_int, C.POINTER(C.c_int), _dp, C.c_double, C.c_int, C.c_double]
        L.shim_close.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def write_reference_xml(path, variant="ocp", exponent=False):
    """The shipped VGP (values of resource/configs/ocp_2d_ex1.xml / mip_2d_ex1.xml, restated in
    etol_b200.workloads) written in the ETOL XML wire format by this test, not copied from the tree."""
    nsteps, nc = (32, 2) if variant == "ocp" else (16, 4)
    fmt = (lambda v: "%.2e" % v) if exponent else (lambda v: "%.2f" % v)
    o = ['<?xml version="1.0" encoding="UTF-8"?>', f'<etol nsteps="{nsteps}" dt="{fmt(0.5)}">',
         ' <states nstates="2" rhorizon="0">']
    for i, (x0, xf) in enumerate(((1.0, 5.0), (2.0, 4.0))):
        o.append(f'  <state name="x{i}" vartype="C" lower="{fmt(0)}" upper="{fmt(7)}" initial="{fmt(x0)}" '
                 f'terminal="{fmt(xf)}" tolerance="{fmt(0.01)}"/>')
    o += [' </states>', f' <controls ncontrols="{nc}" rhorizon="0">']
    for j in range(nc):
        o.append(f'  <control name="u{j}" vartype="C" lower="{fmt(-0.5)}" upper="{fmt(0.5)}"/>')
    o += [' </controls>', f' <exzones nzones="{len(W.REF_BORDERS)}">']
    for z, poly in enumerate(W.REF_BORDERS):
        o.append(f'  <border name="exz{z}" ncorners="{len(poly)}">')
        for c in poly:
            o.append(f'   <corner x="{fmt(c[0])}" y="{fmt(c[1])}" z="{fmt(c[2])}"/>')
        o.append('  </border>')
    o += [' </exzones>', f' <mexzones nzones="{len(W.REF_TRACKS[variant])}">']
    for z, (r, t, x, y) in enumerate(W.REF_TRACKS[variant]):
        o.append(f'  <track name="mexz{z}" radius="{fmt(r)}" nwaypoints="{len(t)}">')
        for w in range(len(t)):
            o.append(f'   <waypoint name="pt{w}" t="{fmt(t[w])}" ndatums="2">')
            o.append(f'    <datum>{fmt(x[w])}</datum>')
            o.append(f'    <datum>{fmt(y[w])}</datum>')
            o.append('   </waypoint>')
        o.append('  </track>')
    o += [' </mexzones>', '</etol>']
    with open(path, "w") as f:
        f.write("\n".join(o) + "\n")
    return path


class Plugin:
    def __init__(self):
        self.L = lib()
        self.h = C.c_void_p(self.L.shim_create())

    def load(self, xml, model=W.SI2D, obstacles=True, tracks=True, batch=1, scaling="none", derivatives="automatic"):
        flags = (1 if obstacles else 0) | (2 if tracks else 0)
        self.L.shim_load(self.h, xml.encode(), model, flags, batch, scaling.encode(), derivatives.encode())
        self.dims = capi.Dims()
        self.L.shim_dims(self.h, C.byref(self.dims))
        return self

    def load_callbacks(self, xml, variant=0):
        """register the example's ecuda::var callbacks (tests/plugin/shim.cpp) and match them"""
        model, flags, why = C.c_int(-1), C.c_int(0), C.create_string_buffer(256)
        ok = self.L.shim_load_callbacks(self.h, xml.encode(), variant, C.byref(model), C.byref(flags), why)
        if ok:
            self.dims = capi.Dims()
            self.L.shim_dims(self.h, C.byref(self.dims))
        return bool(ok), model.value, flags.value, why.value.decode()

    def vgp(self):
        out, dt = (C.c_int * 6)(), C.c_double()
        self.L.shim_vgp(self.h, out, C.byref(dt))
        return dict(nsteps=out[0], nstates=out[1], ncontrols=out[2], nzones=out[3], ntracks=out[4], nparams=out[5],
                    dt=dt.value)

    def bounds(self):
        d = self.dims
        a = [np.zeros(d.nvars), np.zeros(d.nvars), np.zeros(d.ncons), np.zeros(d.ncons), np.zeros(d.nvars),
             np.zeros(d.nvars), np.zeros(d.ncons)]
        self.L.shim_bounds(self.h, *[x.ctypes.data_as(_dp) for x in a])
        return dict(zip(("zl", "zu", "gl", "gu", "guess", "sz", "sg"), a))

    def instance(self, b=0):
        out = np.zeros(self.dims.inst_stride)
        self.L.shim_instance(self.h, b, out.ctypes.data_as(_dp))
        return out

    def edit_instance_and_remesh(self, b, index, value, more_nodes=8, as_setup=False):
        return float(self.L.shim_edit_instance_and_remesh(self.h, b, index, value, more_nodes, int(as_setup)))

    def structure(self):
        d = self.dims
        irow, jcol, grp = np.zeros(d.nnz, np.int32), np.zeros(d.nnz, np.int32), np.zeros(d.nvars, np.int32)
        self.L.shim_structure(self.h, irow.ctypes.data_as(_ip), jcol.ctypes.data_as(_ip), grp.ctypes.data_as(_ip))
        return irow, jcol, grp

    def save_xml(self, path):
        self.L.shim_save_xml(self.h, path.encode())

    # ---- device path
    def setup(self):
        self.L.shim_setup(self.h)

    def evaluate(self, z, batch=1):
        d = self.dims
        z = np.ascontiguousarray(z, dtype=np.float64)
        f, g, jac = np.zeros(batch), np.zeros((batch, d.ncons)), np.zeros((batch, d.nnz))
        rc = self.L.shim_evaluate(self.h, z.ctypes.data_as(_dp), f.ctypes.data_as(_dp), g.ctypes.data_as(_dp),
                                  jac.ctypes.data_as(_dp))
        assert rc == 0
        return f, g, jac

    def solve(self, max_iter=300, print_level=0):
        score, iters, viol = C.c_double(), C.c_int(), C.c_double()
        rc = self.L.shim_solve(self.h, max_iter, print_level, C.byref(score), C.byref(iters), C.byref(viol))
        return rc, score.value, iters.value, viol.value

    def set_hessian(self, mode):
        self.L.shim_set_hessian(self.h, mode.encode())

    def set_mesh(self, mode="automatic", ode_tolerance=1e-4, max_iterations=10):
        self.L.shim_set_mesh(self.h, mode.encode(), ode_tolerance, max_iterations)

    def mesh_history(self):
        nodes, err = (C.c_int * 32)(), np.zeros(32)
        n = self.L.shim_mesh_history(self.h, 32, nodes, err.ctypes.data_as(_dp))
        return [(nodes[i], float(err[i])) for i in range(n)]

    def traj(self, which, width, n):
        out = np.zeros((n, 1 + width))
        got = self.L.shim_traj(self.h, which, out.ctypes.data_as(_dp))
        assert got == n
        return out

    def close(self):
        if self.h:
            self.L.shim_close(self.h)
            self.L.shim_destroy(self.h)
            self.h = None


def next_mesh_size(history, tol=1e-4, initial_increment=10, factor=0.4):
    nodes = (C.c_int * len(history))(*[h[0] for h in history])
    err = np.array([h[1] for h in history], dtype=np.float64)
    return lib().shim_next_mesh_size(len(history), nodes, err.ctypes.data_as(_dp), tol, initial_increment, factor)


def save_csv(path, n, width):
    buf = C.create_string_buffer(4096)
    lib().shim_save_csv(path.encode(), n, width, buf, 4096)
    return buf.value.decode()


def interp(t, tv, ref):
    tv, ref = np.ascontiguousarray(tv, dtype=np.float64), np.ascontiguousarray(ref, dtype=np.float64)
    return lib().shim_interp(float(t), len(tv), tv.ctypes.data_as(_dp), ref.ctypes.data_as(_dp))


def nlp_solve(n, m, zl, zu, gl, gu, irow, jcol, evalfn, z0, max_iter=300, tol=1e-6, print_level=0):
    """built-in interior-point driver with a python evaluation callback evalfn(z, want) -> dict"""
    nnz = len(irow)

    def cb(zp, fp, gp, jp, dp):
        z = np.ctypeslib.as_array(zp, shape=(n,)).copy()
        want = [k for k, p in (("f", fp), ("g", gp), ("jac", jp), ("grad", dp)) if p]
        r = evalfn(z, want)
        if fp:
            fp[0] = float(r["f"])
        if gp:
            np.ctypeslib.as_array(gp, shape=(m,))[:] = r["g"]
        if jp:
            np.ctypeslib.as_array(jp, shape=(nnz,))[:] = r["jac"]
        if dp:
            np.ctypeslib.as_array(dp, shape=(n,))[:] = r["grad"]
        return 0

    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (zl, zu, gl, gu)]
    irow, jcol = np.ascontiguousarray(irow, np.int32), np.ascontiguousarray(jcol, np.int32)
    z = np.ascontiguousarray(z0, dtype=np.float64).copy()
    res = np.zeros(3)
    rc = lib().shim_nlp_solve(n, m, nnz, *[a.ctypes.data_as(_dp) for a in arrs], irow.ctypes.data_as(_ip),
                              jcol.ctypes.data_as(_ip), EVAL_CB(cb), max_iter, tol, print_level,
                              z.ctypes.data_as(_dp), res.ctypes.data_as(_dp))
    return rc, z, dict(iterations=int(res[0]), objective=res[1], max_violation=res[2])


def gen_region(xy):
    """convex partition of a polygon -> list of (lower [n][2], upper [m][2], lower slopes, upper slopes)"""
    xy = np.ascontiguousarray(xy, dtype=np.float64)
    out = np.zeros(64 * len(xy) + 64)
    n = lib().shim_gen_region(xy.ctypes.data_as(_dp), len(xy), out.ctypes.data_as(_dp), len(out))
    assert n >= 0
    pieces, w = [], 0
    for _ in range(n):
        nl, nu = int(out[w]), int(out[w + 1])
        w += 2
        lo = out[w:w + 2 * nl].reshape(nl, 2).copy(); w += 2 * nl
        up = out[w:w + 2 * nu].reshape(nu, 2).copy(); w += 2 * nu
        sl = out[w:w + nl - 1].copy(); w += nl - 1
        su = out[w:w + nu - 1].copy(); w += nu - 1
        pieces.append((lo, up, sl, su))
    return pieces
