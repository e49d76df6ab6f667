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


EXT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c0_ext_golden.npz")


def _user_workload():
    return W.unicycle(batch=1, nnodes=17, ncyl=2, ntracks=1)


def test_oracle_reproduces_extended_golden():
    """Hessian of the Lagrangian, discretisation error and resampling on C0; a user model (unicycle tape)"""
    gold = np.load(EXT)
    wl = W.reference_vgp("ocp")
    assert np.array_equal(wl.x, gold["x"])
    o = ob.Oracle(wl)
    hi, hj = ob.hess_structure(o)
    assert np.array_equal(hi, gold["hess_irow"]) and np.array_equal(hj, gold["hess_jcol"])
    assert rel_err(ob.eval_hess(o, wl.x, gold["sigma"], gold["lam"]), gold["hess"]) <= TOL_JAC
    assert rel_err(ob.ode_error(o, wl, wl.x), gold["ode_error"]) <= TOL_VALUE
    assert rel_err(ob.resample(o, wl, wl.x, [41]), gold["resample_41"]) <= TOL_VALUE
    uw = _user_workload()
    assert np.array_equal(uw.x, gold["user_x"]), "user workload generator drifted from the committed fixture"
    uo = ob.Oracle(uw)
    fd = uo.eval(uw.x, want=("f", "g", "jac"), jac_mode=1)
    ex = uo.eval(uw.x, want=("jac",), jac_mode=0)
    assert rel_err(fd["f"], gold["user_f"]) <= TOL_VALUE and rel_err(fd["g"], gold["user_g"]) <= TOL_VALUE
    assert rel_err(fd["jac"], gold["user_jac_fd"]) <= TOL_JAC and rel_err(ex["jac"], gold["user_jac_exact"]) <= TOL_JAC


def test_stepped_user_model_reproduces_golden():
    """the generated Model<ECUDA_MODEL_USER> source, compiled into the emulator, against the committed vectors"""
    gold = np.load(EXT)
    uw = _user_workload()
    for mode, key in ((1, "user_jac_fd"), (0, "user_jac_exact")):
        got = eb.emu_eval(uw, uw.x, jac_mode=mode, nthr=256)
        assert rel_err(got["f"], gold["user_f"]) <= TOL_VALUE and rel_err(got["g"], gold["user_g"]) <= TOL_VALUE
        assert rel_err(got["jac"], gold[key]) <= TOL_JAC
