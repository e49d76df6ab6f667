"""The oracle and the CPU-stepped kernel phases against the committed golden fixture."""
import os

import numpy as np

import emu_binding as eb
import oracle_binding as ob
from conftest import TOL_JAC, TOL_VALUE, rel_err
from etol_b200 import workloads as W

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c0_ocp_golden.npz")


def test_oracle_reproduces_golden():
    gold = np.load(GOLD)
    wl = W.reference_vgp("ocp")
    assert np.array_equal(wl.x, gold["x"])
    o = ob.Oracle(wl)
    fd = o.eval(wl.x, want=("f", "g", "jac", "grad"), jac_mode=1)
    ex = o.eval(wl.x, want=("jac",), jac_mode=0)
    assert rel_err(fd["f"], gold["f"]) <= TOL_VALUE and rel_err(fd["g"], gold["g"]) <= TOL_VALUE
    assert rel_err(fd["jac"], gold["jac_fd"]) <= TOL_JAC and rel_err(ex["jac"], gold["jac_exact"]) <= TOL_JAC
    irow, jcol, grp = o.structure()
    assert np.array_equal(irow, gold["irow"]) and np.array_equal(jcol, gold["jcol"])
    assert np.array_equal(grp, gold["group_of_col"])


def test_stepped_kernel_phases_reproduce_golden():
    gold = np.load(GOLD)
    wl = W.reference_vgp("ocp")
    got = eb.emu_eval(wl, wl.x, jac_mode=1)
    assert rel_err(got["f"], gold["f"]) <= TOL_VALUE and rel_err(got["g"], gold["g"]) <= TOL_VALUE
    assert rel_err(got["jac"], gold["jac_fd"]) <= TOL_JAC
