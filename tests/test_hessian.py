"""Hessian of the Lagrangian (the reference asks for hessian = "exact", ePSOPT.cpp:65). Without a GPU: the
structure (host helper of libecuda.so against the oracle) and the oracle's second-order forward mode
against central differences of its own exact Lagrangian gradient. The kernel is compared with the oracle in
tests/test_gpu_parity.py."""
import numpy as np
import pytest

import oracle_binding as ob
from etol_b200 import capi, workloads as W

CASES = {
    "C0-ocp": lambda: W.reference_vgp("ocp", batch=2, jitter=0.02),
    "C0-mip-max": lambda: W.reference_vgp("mip", maximize=True),
    "pm3d-scaled": lambda: W.pm3d(batch=2, nnodes=9, ncyl=3, scaled=True),
    "fw6": lambda: W.fw6(batch=1, nnodes=7, ncyl=2, scaled=True),
    "multiphase": lambda: W.pm3d_multiphase(batch=1, nphases=2, nnodes=6, ncyl=1),
    "user-unicycle-tracks": lambda: W.unicycle(batch=1, nnodes=8, ncyl=2, ntracks=1),
    "user-dragmass": lambda: W.dragmass(batch=1, nnodes=8, ncyl=1, scaled=True),
    # round 2: dynamics / cost that read t, and traced path rows
    "user-gust": lambda: W.gust(batch=1, nnodes=8, ncyl=1, ntracks=1),
    "user-zone-gust": lambda: W.zone(batch=1, timedep=True, nnodes=7, ncyl=1, scaled=True),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_structure(name):
    wl = CASES[name]()
    o = ob.Oracle(wl)
    irow, jcol = capi.host_hess_structure(wl)
    oi, oj = ob.hess_structure(o)
    assert np.array_equal(irow, oi) and np.array_equal(jcol, oj)
    ns, nc = wl.ns, wl.nc
    per_node = nc * (nc + 1) // 2 + nc * (ns + 2) + ns * (ns + 1) // 2 + 2 * ns
    assert len(irow) == sum(n * per_node + 3 for n in wl.nnodes)
    assert (irow >= jcol).all()                                           # lower triangle
    key = jcol.astype(np.int64) * (wl.nvars + 1) + irow
    assert (np.diff(key) > 0).all()                                       # sorted by (column, row), no duplicates


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_hessian_is_the_derivative_of_the_lagrangian_gradient(name):
    wl = CASES[name]()
    o = ob.Oracle(wl)
    rng = np.random.default_rng(11)
    B, n, m = wl.batch, wl.nvars, wl.ncons
    lam = rng.normal(size=(B, m))
    sigma = rng.uniform(0.5, 2.0, size=B)
    irow, jcol = ob.hess_structure(o)
    jr, jc, _ = o.structure()
    base = wl.index_base
    H = ob.eval_hess(o, wl.x, sigma, lam)

    def grad_lagrangian(x):
        r = o.eval(x, want=("jac", "grad"), jac_mode=W.JAC_EXACT)
        out = sigma[:, None] * r["grad"]
        for b in range(B):
            np.add.at(out[b], jc - base, r["jac"][b] * lam[b, jr - base])
        return out

    for trial in range(3):
        v = rng.normal(size=(B, n))
        eps = 1e-6 * (1.0 + np.abs(wl.x))
        fd = (grad_lagrangian(wl.x + eps * v) - grad_lagrangian(wl.x - eps * v)) / 2.0
        Hv = np.zeros((B, n))
        for b in range(B):
            dense = np.zeros((n, n))
            dense[irow - base, jcol - base] = H[b]
            dense = dense + dense.T - np.diag(np.diag(dense))
            Hv[b] = dense @ (eps[b] * v[b])
        scale = np.abs(fd).max() + np.abs(Hv).max()
        assert np.abs(fd - Hv).max() <= 2e-5 * scale, name
