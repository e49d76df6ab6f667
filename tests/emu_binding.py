"""ctypes binding of tests/emu/libecuda_emu.so: the kernel phases stepped on the CPU (test-only)."""
import ctypes as C
import os
import subprocess

import numpy as np

from etol_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "tests", "emu")
LIB = os.path.join(EMU_DIR, "libecuda_emu.so")
_dp = C.POINTER(C.c_double)
_lib = None


def build():
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-mfma", "-Wno-unknown-pragmas",
                    "-shared", "-o", LIB, os.path.join(EMU_DIR, "emu.cpp"),
                    os.path.join(ROOT, "etol_b200", "csrc", "ecuda_host.cpp")], check=True)


def lib():
    global _lib
    if _lib is None:
        srcs = [os.path.join(EMU_DIR, "emu.cpp")] + [
            os.path.join(ROOT, "etol_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "etol_b200", "csrc"))
            if f.endswith((".cuh", ".hpp", ".cpp"))] + [os.path.join(ROOT, "include", "ecuda_detmath.h")]
        newest = max(os.path.getmtime(s) for s in srcs)
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
            build()
        _lib = C.CDLL(LIB)
        _lib.emu_eval.argtypes = [C.POINTER(capi.ProblemDesc), _dp, _dp, C.c_double, _dp, _dp, _dp, _dp, _dp, _dp,
                                  C.c_int, C.c_int, C.c_int]
    return _lib


def emu_eval(wl, x, want=("f", "g", "jac"), jac_mode=1, nthr=64, generic=False, variant="rows"):
    """variant: which specialised kernel the emulator steps when the problem qualifies -- "rows"
    (k_eval_rows, the default), "columns" (k_eval_fast) or "image" (k_eval_image, exact mode)"""
    lib().emu_use_rows(1 if variant == "rows" else 0)
    lib().emu_use_image(1 if variant == "image" else 0)
    dims = capi.host_dims(wl)
    inst = capi.pack_instances(wl, dims)
    x = np.ascontiguousarray(x, dtype=np.float64).reshape(wl.batch, dims.nvars)
    f = np.zeros(wl.batch) if "f" in want else None
    g = np.zeros((wl.batch, dims.ncons)) if "g" in want else None
    jac = np.full((wl.batch, dims.nnz), np.nan) if "jac" in want else None
    grad = np.full((wl.batch, dims.nvars), np.nan) if "grad" in want else None
    p = lambda a: None if a is None else a.ctypes.data_as(_dp)
    sz = None if wl.sz is None else np.ascontiguousarray(wl.sz, dtype=np.float64)
    sg = None if wl.sg is None else np.ascontiguousarray(wl.sg, dtype=np.float64)
    desc = capi.make_desc(wl)
    rc = lib().emu_eval(C.byref(desc), p(sz), p(sg), float(wl.sf), p(inst), p(x), p(f), p(g), p(jac), p(grad),
                        jac_mode, nthr, int(generic))
    if rc != 0:
        raise RuntimeError("emu_eval failed")
    return dict(f=f, g=g, jac=jac, grad=grad)
