"""src/eCUDA/ecuda_nlp_ipopt.cpp (the IPOPT TNLP adapter eCUDA::solve() uses where IPOPT exists, reference:
src/ePSOPT/ePSOPT.cpp:62,84) is compiled against a stub of IpStdCInterface.h and driven once, so that it cannot rot
behind its #ifdef in an image without IPOPT. The stub "solver" only queries the structures and evaluates every
callback at the starting point: this checks the adapter's plumbing, not an optimisation."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib(tmp_path):
    out = str(tmp_path / "libipopt_adapter_test.so")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-std=c++17", "-O1", "-fPIC", "-Wall", "-Werror", "-DECUDA_HAVE_IPOPT", "-shared", "-o", out,
                    "-I" + os.path.join(ROOT, "tests", "ipopt_stub"), "-I" + os.path.join(ROOT, "src", "eCUDA"),
                    os.path.join(ROOT, "src", "eCUDA", "ecuda_nlp_ipopt.cpp"),
                    os.path.join(ROOT, "tests", "ipopt_stub", "stub_ipopt.cpp")], check=True)
    L = C.CDLL(out)
    L.stub_run.argtypes = [C.c_int, C.POINTER(C.c_double), C.c_int]
    return L


def test_ipopt_adapter_compiles_and_serves_every_callback(tmp_path):
    L = _lib(tmp_path)
    for with_h in (0, 1):
        buf = np.zeros(64)
        n = L.stub_run(with_h, buf.ctypes.data_as(C.POINTER(C.c_double)), 64)
        v = buf[:n]
        rc, obj, viol, have, exact, max_iter, xu1 = v[:7]
        assert rc == 0 and have == 1 and max_iter == 77 and xu1 == 2e19
        assert obj == 13.0 and viol == 0.0          # f(2,3) = 13, g = 6 >= 1
        assert exact == with_h                       # exact Hessian when the problem carries eval_h, else limited-memory
        seen = v[7:]
        assert list(seen[:3]) == [13.0, 4.0, 6.0]    # objective, gradient
        assert list(seen[3:9]) == [0, 0, 3.0, 0, 1, 2.0]  # Jacobian triplets (row, col, value)
        if with_h:
            assert list(seen[9:18]) == [0, 0, 2.0, 1, 0, 0.5, 1, 1, 2.0]  # lower triangle, lambda = 0.5, sigma = 1
        else:
            assert len(seen) == 9
