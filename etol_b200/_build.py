"""In-tree build of the native pieces (no JIT cache: the .so files travel with the repo snapshot).

  etol_b200/csrc/libecuda.so   CUDA kernels + C ABI, nvcc -gencode arch=compute_100a,code=sm_100a
  oracle/_build/liboracle.so   CPU oracle (test infrastructure), g++
  oracle/_ref/libetol_ref.so   the reference's ePSOPT.cpp + example, unmodified, against a stub psopt.h (test infrastructure)
  tests/emu/libecuda_emu.so    test-only host stepping of the kernel phases, g++
  build/libetol_ecuda.so       C++ plugin layer (TrajectoryOptimizer core-lite + eCUDA), g++ (when present)
"""
import os
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "etol_b200", "csrc")
LIBECUDA = os.path.join(CSRC, "libecuda.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # the canonical operation order of DESIGN.md section 3: no implicit fma on either side
    "-fmad=false", "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math",
]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libecuda.so cannot be built (eCUDA has no CPU fallback)")


def _cxx():
    return "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def cuda_sources():
    deps = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".hpp", ".cpp"))]
    deps += [os.path.join(ROOT, "include", f) for f in ("ecuda.h", "ecuda_detmath.h")]
    return deps


def build_libecuda(force=False, verbose=False):
    """nvcc sm_100a. ecuda_api.cu is compiled as two translation units (the second one, ecuda_api_rowsn.cu, holds the
    instantiations of the N-specialised kernel family) next to the three host files, all in parallel, then linked."""
    srcs = [os.path.join(CSRC, f) for f in ("ecuda_api.cu", "ecuda_api_rowsn.cu", "ecuda_host.cpp", "ecuda_usermodel.cpp",
                                            "ecuda_mesh.cpp")]
    if not force and not _stale(LIBECUDA, cuda_sources()):
        return LIBECUDA
    from concurrent.futures import ThreadPoolExecutor
    objdir = os.path.join(ROOT, "build", "obj")
    os.makedirs(objdir, exist_ok=True)
    base = [_nvcc(), "-ccbin", _cxx()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else [])

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src) + ".o")
        subprocess.run(base + ["-c", "-o", obj, src], check=True)
        return obj

    with ThreadPoolExecutor(len(srcs)) as ex:
        objs = list(ex.map(compile_one, srcs))
    subprocess.run(base + ["-shared", "-o", LIBECUDA] + objs + ["-ldl"], check=True)
    return LIBECUDA


def build_oracle():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    # oracle/_ref: the reference's own ePSOPT.cpp + example compiled unmodified against oracle/refstub/psopt.h.
    # Only where the reference tree exists (this container); the GPU box uses the built file that travelled.
    if os.path.isdir("/root/reference/src/ePSOPT"):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True)


def build_emu():
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import emu_binding
    emu_binding.lib()


def build_plugin():
    mk = os.path.join(ROOT, "src", "Makefile")
    if os.path.exists(mk):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "src")], check=True)
