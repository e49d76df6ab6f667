"""ctypes binding of oracle/_ref/libetol_ref.so: the reference's own ePSOPT.cpp + etol_psopt_example1.cpp, compiled
unmodified against oracle/refstub/psopt.h (recipe: `make -C oracle ref`, needs /root/reference). Test infrastructure."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "libetol_ref.so")
REF_XML = "/root/reference/resource/configs/ocp_2d_ex1.xml"
_dp = C.POINTER(C.c_double)
_lib = None


def available():
    if os.path.isdir("/root/reference/src/ePSOPT"):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True)
    return os.path.exists(LIB)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB)
        L.ref_open.restype = C.c_void_p
        L.ref_open.argtypes = [C.c_char_p, C.c_int]
        L.ref_close.argtypes = [C.c_void_p]
        L.ref_dims.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.ref_bounds.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, _dp]
        L.ref_algorithm.argtypes = [C.c_void_p, C.c_char_p, C.c_int, _dp]
        L.ref_dae.argtypes = [C.c_void_p, _dp, _dp, C.c_double, _dp, _dp, _dp, _dp]
        L.ref_integrand_cost.restype = C.c_double
        L.ref_integrand_cost.argtypes = [C.c_void_p, _dp, _dp, C.c_double, _dp]
        L.ref_endpoint_cost.restype = C.c_double
        L.ref_endpoint_cost.argtypes = [C.c_void_p, _dp, _dp, C.c_double, C.c_double]
        L.ref_events.argtypes = [C.c_void_p, _dp, _dp, C.c_double, C.c_double, _dp]
        L.ref_guess_time.argtypes = [C.c_void_p, _dp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(_dp)


class Reference:
    """ETOL::ePSOPT set up the way the reference example's main() does it (etol_psopt_example1.cpp:41-66)."""

    def __init__(self, xml, maximize=False):
        self.L = lib()
        self.h = C.c_void_p(self.L.ref_open(xml.encode(), int(maximize)))
        d = (C.c_int * 5)()
        self.L.ref_dims(self.h, d)
        self.ns, self.nc, self.ne, self.npath, self.nodes = list(d)
        self.ndir = self.ns + self.nc + 1

    def close(self):
        if self.h:
            self.L.ref_close(self.h)
            self.h = None

    def bounds(self):
        xs, us, ev, pa, tm = (np.zeros(2 * self.ns), np.zeros(2 * self.nc), np.zeros(2 * self.ne),
                              np.zeros(2 * self.npath), np.zeros(4))
        self.L.ref_bounds(self.h, _p(xs), _p(us), _p(ev), _p(pa), _p(tm))
        return dict(states=xs.reshape(2, -1), controls=us.reshape(2, -1), events=ev.reshape(2, -1),
                    path=pa.reshape(2, -1), t0=tm[:2], tf=tm[2:])

    def algorithm(self):
        buf, nums = C.create_string_buffer(256), np.zeros(5)
        self.L.ref_algorithm(self.h, buf, 256, _p(nums))
        return buf.value.decode().split("|"), nums

    def dae(self, x, u, t):
        """ePSOPT::dae at one node: (derivatives[ns], path[npath], d_derivatives[ns][ndir], d_path[npath][ndir]);
        tangent directions are the states, then the controls, then t"""
        x, u = np.ascontiguousarray(x, dtype=np.float64), np.ascontiguousarray(u, dtype=np.float64)
        f, p = np.zeros(self.ns), np.zeros(max(self.npath, 1))
        df, dp = np.zeros((self.ns, self.ndir)), np.zeros((max(self.npath, 1), self.ndir))
        rc = self.L.ref_dae(self.h, _p(x), _p(u), float(t), _p(f), _p(p), _p(df), _p(dp))
        assert rc == 0
        return f, p[:self.npath], df, dp[:self.npath]

    def integrand_cost(self, x, u, t):
        x, u = np.ascontiguousarray(x, dtype=np.float64), np.ascontiguousarray(u, dtype=np.float64)
        d = np.zeros(self.ndir)
        return self.L.ref_integrand_cost(self.h, _p(x), _p(u), float(t), _p(d)), d

    def endpoint_cost(self, x0, xf, t0, tf):
        x0, xf = np.ascontiguousarray(x0, dtype=np.float64), np.ascontiguousarray(xf, dtype=np.float64)
        return self.L.ref_endpoint_cost(self.h, _p(x0), _p(xf), float(t0), float(tf))

    def events(self, x0, xf, t0, tf):
        x0, xf = np.ascontiguousarray(x0, dtype=np.float64), np.ascontiguousarray(xf, dtype=np.float64)
        e = np.zeros(self.ne)
        self.L.ref_events(self.h, _p(x0), _p(xf), float(t0), float(tf), _p(e))
        return e

    def guess_time(self):
        t = np.zeros(self.nodes)
        self.L.ref_guess_time(self.h, _p(t))
        return t


def node_inputs(wl, x, b=0):
    """states[N][ns], controls[N][nc], t0, tf of instance b of a single-phase workload"""
    N, ns, nc = wl.nnodes[0], wl.ns, wl.nc
    z = np.asarray(x, dtype=np.float64).reshape(wl.batch, -1)[b]
    X = np.array([[z[wl.ix(0, k, i)] for i in range(ns)] for k in range(N)])
    U = np.array([[z[wl.iu(0, k, j)] for j in range(nc)] for k in range(N)])
    return X, U, z[wl.it0(0)], z[wl.itf(0)]
