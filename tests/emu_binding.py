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


def build(out=LIB, user_header=None):
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    extra = [] if user_header is None else ['-DECUDA_USER_MODEL_HEADER="%s"' % user_header]
    tmp = "%s.%d.tmp" % (out, os.getpid())  # atomic: parallel test workers may build the same library
    subprocess.run([cxx, "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-mfma", "-Wno-unknown-pragmas"] + extra +
                   ["-shared", "-o", tmp, os.path.join(EMU_DIR, "emu.cpp"),
                    os.path.join(ROOT, "etol_b200", "csrc", "ecuda_host.cpp"),
                    os.path.join(ROOT, "etol_b200", "csrc", "ecuda_usermodel.cpp"), "-ldl"], check=True)
    os.replace(tmp, out)


def _sources():
    return [os.path.join(EMU_DIR, "emu.cpp")] + [
        os.path.join(ROOT, "etol_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "etol_b200", "csrc"))
        if f.endswith((".cuh", ".hpp", ".cpp"))] + [os.path.join(ROOT, "include", "ecuda_detmath.h"),
                                                    os.path.join(ROOT, "include", "ecuda.h")]


_user_libs = {}


def user_lib(tape):
    """The emulator compiled for one user model: the source libecuda.so generates for the tape is built into a
    g++ copy of the kernel phases (the very text NVRTC compiles for the GPU). Returns (library, model id in it)."""
    import hashlib
    key = tape.key()
    if key in _user_libs:
        return _user_libs[key]
    src = capi.user_model_source(capi.register_user_model(tape))
    tag = hashlib.sha1(src.encode()).hexdigest()[:12]
    udir = os.path.join(EMU_DIR, "_user")
    os.makedirs(udir, exist_ok=True)
    hdr, out = os.path.join(udir, tag + ".cuh"), os.path.join(udir, "libecuda_emu_" + tag + ".so")
    if not os.path.exists(hdr) or open(hdr).read() != src:
        with open(hdr, "w") as fh:
            fh.write(src)
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(s) for s in _sources() + [hdr]):
        build(out, hdr)
    L = C.CDLL(out)
    L.emu_eval.argtypes = lib().emu_eval.argtypes
    L.ecuda_register_user_model.argtypes = capi.lib().ecuda_register_user_model.argtypes
    nodes = (capi.TapeNode * len(tape.nodes))()
    for i, (op, a, b, imm) in enumerate(tape.nodes):
        nodes[i].op, nodes[i].a, nodes[i].b, nodes[i].imm = op, a, b, imm
    um = capi.UserModel()
    um.nstates, um.ncontrols, um.static_kind, um.nnodes = tape.ns, tape.nc, tape.static_kind, len(tape.nodes)
    um.nodes = nodes
    for i, v in enumerate(tape.f_out):
        um.f_out[i] = v
    um.cost_out = tape.cost_out
    mid = C.c_int32(-1)
    rows = np.array(getattr(tape, "row_out", []), dtype=np.int32)
    if rows.size:
        L.ecuda_register_user_model_rows.argtypes = capi.lib().ecuda_register_user_model_rows.argtypes
        assert L.ecuda_register_user_model_rows(C.byref(um), int(rows.size), rows.ctypes.data_as(C.POINTER(C.c_int32)),
                                                C.byref(mid), None, 0) == 0
    else:
        assert L.ecuda_register_user_model(C.byref(um), C.byref(mid), None, 0) == 0
    _user_libs[key] = (L, mid.value)
    return _user_libs[key]


def lib():
    global _lib
    if _lib is None:
        newest = max(os.path.getmtime(s) for s in _sources())
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
            build()
        _lib = C.CDLL(LIB)
        _lib.emu_eval.argtypes = [C.POINTER(capi.ProblemDesc), _dp, _dp, C.c_double, _dp, _dp, _dp, _dp, _dp, _dp,
                                  C.c_int, C.c_int, C.c_int]
    return _lib


def emu_eval(wl, x, want=("f", "g", "jac"), jac_mode=1, nthr=64, generic=False, variant="rows"):
    """variant: which specialised kernel the emulator steps when the problem qualifies -- "rows"
    (k_eval_rows), "rowsn" (the N-specialised k_rows_n_* where an instantiation exists, the product's default),
    "columns" (k_eval_fast) or "image" (k_eval_image, exact mode)"""
    L, model = (lib(), wl.model) if getattr(wl, "tape", None) is None else user_lib(wl.tape)
    L.emu_use_rows(1 if variant in ("rows", "rowsn") else 0)
    L.emu_use_rowsn(1 if variant == "rowsn" else 0)
    L.emu_use_image(1 if variant == "image" else 0)
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
    desc.model = model
    rc = L.emu_eval(C.byref(desc), p(sz), p(sg), float(wl.sf), p(inst), p(x), p(f), p(g), p(jac), p(grad),
                        jac_mode, nthr, int(generic))
    if rc != 0:
        raise RuntimeError("emu_eval failed")
    return dict(f=f, g=g, jac=jac, grad=grad)
