"""The C-ABI library loads and exports every symbol include/ecuda.h declares; without a GPU the
compute entry points fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from etol_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "ecuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ecuda_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    L = capi.lib()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), f"libecuda.so does not export {n}"
    assert sorted(capi.ABI_SYMBOLS) == names
    assert L.ecuda_abi_version() == 1


def test_library_is_self_contained():
    """the product library must not link or load the oracle or the test emulator"""
    import subprocess
    out = subprocess.run(["ldd", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "emu" not in out
    syms = subprocess.run(["nm", "-D", "--defined-only", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle_" not in syms and "emu_eval" not in syms


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_have_gpu(), reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_gpu():
    h = C.c_void_p()
    rc = capi.lib().ecuda_create(0, C.byref(h))
    assert rc == -3 and not h.value
    msg = capi.lib().ecuda_last_error(None).decode()
    assert "no CPU fallback" in msg
    from etol_b200 import workloads as W
    with pytest.raises(capi.EcudaError):
        capi.Evaluator(W.reference_vgp("ocp"))


def test_null_handle_is_an_argument_error():
    L = capi.lib()
    assert L.ecuda_get_dims(None, C.byref(capi.Dims())) == -1
    assert L.ecuda_eval(None, None, None, None, None, 0, 0, None) == -1
    assert L.ecuda_destroy(None) == 0
