"""CPU stepping of the kernel phases (tests/emu) against the oracle. This is NOT the parity gate
(that is tests/test_gpu_parity.py on a B200, through the C ABI); it checks, where there is no GPU,
that the kernel source computes every row and every triplet with the canonical operation order."""
import numpy as np
import pytest

import emu_binding as eb
import oracle_binding as ob
from conftest import TOL_JAC, TOL_VALUE, rel_err
from etol_b200 import workloads as W

CASES = {
    "C0-ocp": lambda: W.reference_vgp("ocp", batch=3, jitter=0.02),
    "C0-mip": lambda: W.reference_vgp("mip", batch=2, jitter=0.01),
    "C0-ocp-cheb-max": lambda: W.reference_vgp("ocp", collocation=W.CHEBYSHEV, maximize=True),
    "C0-ocp-deps": lambda: W.reference_vgp("ocp", pattern_mode=W.MODEL_DEPS),
    "C2-pm3d": lambda: W.pm3d(batch=3),
    "C2-pm3d-scaled-deps": lambda: W.pm3d(batch=2, scaled=True, pattern_mode=W.MODEL_DEPS),
    "C3-fw6-small": lambda: W.fw6(batch=2, nnodes=21, ncyl=5, scaled=True),
    "C3-fw6-N200": lambda: W.fw6(batch=1),
    "C4-multiphase": lambda: W.pm3d_multiphase(batch=2, scaled=True),
    "C4-multiphase-ragged": lambda: W.pm3d_multiphase(batch=1, nphases=4, nnodes=7, ncyl=1),
    "pm3d-N2": lambda: W.pm3d(batch=1, nnodes=2, ncyl=1),
    "pm3d-no-obstacles": lambda: W.pm3d(batch=2, nnodes=9, ncyl=0),
    # user models: the generated Model<ECUDA_MODEL_USER> source compiled into a copy of the emulator
    "user-unicycle-tracks": lambda: W.unicycle(batch=2, ntracks=1, scaled=True),
    "user-dragmass-deps": lambda: W.dragmass(batch=2, ntracks=2, pattern_mode=W.MODEL_DEPS),
    "user-dragmass-N12": lambda: W.dragmass(batch=2, nnodes=12, ncyl=2, collocation=W.CHEBYSHEV),
    # dynamics and running cost that read t (the `k` argument of the ePSOPT callbacks)
    "user-gust-tracks": lambda: W.gust(batch=2, ntracks=1, scaled=True),
    "user-gust-N12-max": lambda: W.gust(batch=2, nnodes=12, ncyl=2, collocation=W.CHEBYSHEV, maximize=True),
    # traced path constraints (rows that are none of the built-in zone rows), with and without moving zones
    "user-zone-tracks": lambda: W.zone(batch=2, ntracks=2, scaled=True),
    "user-zone-gust-N12": lambda: W.zone(batch=2, timedep=True, nnodes=12, ncyl=0, pattern_mode=W.MODEL_DEPS),
}


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("nthr,generic,variant", [(32, False, "rows"), (256, False, "rows"), (256, False, "rowsn"),
                                                  (256, False, "columns"),
                                                  (256, False, "image"), (256, True, "rows")])
def test_phases_match_oracle(name, nthr, generic, variant):
    wl = CASES[name]()
    if name == "C3-fw6-N200" and (nthr == 32 or generic or variant != "rows"):
        pytest.skip("one thread count is enough for the large case")
    if name.startswith("user-") and variant not in ("rows", "rowsn"):
        pytest.skip("user models are compiled for the row-owner and the generic kernels only")
    o = ob.Oracle(wl)
    style = 1 if name == "C3-fw6-N200" else 0
    for mode in (W.JAC_FD, W.JAC_EXACT):
        ref = o.eval(wl.x, want=("f", "g", "jac", "grad"), jac_mode=mode, style=style, nthreads=4)
        got = eb.emu_eval(wl, wl.x, want=("f", "g", "jac", "grad"), jac_mode=mode, nthr=nthr, generic=generic,
                          variant=variant)
        assert rel_err(got["f"], ref["f"]) <= TOL_VALUE
        assert rel_err(got["g"], ref["g"]) <= TOL_VALUE
        assert rel_err(got["grad"], ref["grad"]) <= TOL_VALUE
        assert not np.isnan(got["jac"]).any(), "a triplet was never written"
        assert rel_err(got["jac"], ref["jac"]) <= TOL_JAC
        if mode == W.JAC_FD:  # same operation sequence => identical bits
            assert np.array_equal(got["g"], ref["g"]) and np.array_equal(got["jac"], ref["jac"])
