"""Pins the CPU oracle against every known answer the reference's in-tree formulas give
 and against closed-form identities of the pseudospectral method.
The reference ships no tests or golden vectors: parity is otherwise unpinned."""
import json
import os

import numpy as np
import pytest

import oracle_binding as ob
from conftest import rel_err
from etol_b200 import workloads as W

HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "appendix_b_kats.json")))


def test_edge_geometry_kats_bitwise():
    i = 0
    for poly in W.REF_BORDERS:
        n = len(poly)
        for e in range(n):
            g = ob.edge_geometry(poly[e], poly[(e + 1) % n])
            assert list(g[:4]) == KAT["edge_params_xc_yc_radsq_tt"][i], f"edge {i}"
            assert g[4] == g[2] and g[5] == .2 * g[2]
            i += 1
    assert i == 9


def _g_at(o, wl, xy, N):
    """constraint vector with every node's position states set to xy"""
    x = wl.x.copy()
    for k in range(N):
        x[0, wl.ix(0, k, 0)], x[0, wl.ix(0, k, 1)] = xy
    return o.eval(x, want=("g",))["g"][0]


def test_path_rows_kats():
    wl = W.reference_vgp("ocp")
    o = ob.Oracle(wl)
    N = 33
    tau, w, D = o.collocation(0, N)
    t = 8.0 * tau + 8.0  # t0 = 0, tf = 16
    p0 = 2 * N + 4
    g = _g_at(o, wl, (3.0, 3.0), N)
    k8 = 16
    assert t[k8] == 8.0
    rows = g[p0 + k8 * 11:p0 + (k8 + 1) * 11]
    assert rel_err(rows[:9], KAT["obs_at_3_3"]) <= 1e-13
    assert rows[4] > 0.0  # edge 4 is violated at (3,3)
    # track 0 reaches x = 1.51 + 0.49*t/32; the reference table is defined over t in [0,32]
    assert rel_err(rows[9:], KAT["saa_at_3_3_t8"]) <= 1e-13
    g = _g_at(o, wl, (1.0, 2.0), N)
    rows0 = g[p0:p0 + 11]
    assert t[0] == 0.0
    assert rel_err(rows0[:2], KAT["obs_at_1_2_first2"]) <= 1e-13
    assert rel_err(rows0[9:], KAT["saa_at_1_2_t0"]) <= 1e-13


@pytest.mark.parametrize("name,mk", [
    ("ocp", lambda: W.reference_vgp("ocp")), ("mip", lambda: W.reference_vgp("mip")),
    ("pm3d_N40_8cyl", lambda: W.pm3d(batch=1)), ("fw6_N200_64cyl", lambda: W.fw6(batch=1)),
    ("pm3d_3x30_8cyl", lambda: W.pm3d_multiphase(batch=1))])
def test_dimension_kats(name, mk):
    wl = mk()
    o = ob.Oracle(wl)
    assert [o.nvars, o.ncons, o.nnz, o.ngroups] == KAT["dims"][name]
    assert (wl.nvars, wl.ncons) == (o.nvars, o.ncons)
    # closed forms of SURVEY.md section 8(d) for single-phase problems
    if wl.nphases == 1:
        ns, nc, N, npth = wl.ns, wl.nc, wl.nnodes[0], wl.npath[0]
        assert o.nvars == (ns + nc) * N + 2
        assert o.ncons == ns * N + 2 * ns + npth * N + 1
        assert o.ngroups == ns * N + nc + 2


@pytest.mark.parametrize("N", [2, 3, 4, 9, 17, 33, 40, 64])
def test_legendre_identities(N):
    tau, w, D = ob.make_collocation(0, N)
    No = N - 1
    assert tau[0] == -1.0 and tau[-1] == 1.0 and np.all(np.diff(tau) > 0)
    assert np.array_equal(tau, -tau[::-1])
    assert D[0, 0] == -No * (No + 1) / 4.0 and D[-1, -1] == No * (No + 1) / 4.0
    assert abs(w.sum() - 2.0) < 1e-13
    assert np.abs(D @ np.ones(N)).max() < 1e-10 * max(1, No) ** 2
    # exact differentiation / integration of polynomials up to the order of the rule
    for deg in range(0, min(No, 12) + 1):
        p = tau ** deg
        dp = deg * tau ** max(deg - 1, 0) if deg > 0 else np.zeros(N)
        assert np.abs(D @ p - dp).max() < 1e-9 * max(1, No) ** 2
    for deg in range(0, min(2 * No - 1, 15) + 1):
        exact = 0.0 if deg % 2 else 2.0 / (deg + 1)
        assert abs(w @ tau ** deg - exact) < 1e-12


@pytest.mark.parametrize("N", [2, 3, 5, 8, 17, 40])
def test_chebyshev_identities(N):
    tau, w, D = ob.make_collocation(1, N)
    No = N - 1
    assert np.allclose(tau, -np.cos(np.pi * np.arange(N) / No), atol=1e-15)
    assert D[0, 0] == -(2.0 * No * No + 1.0) / 6.0
    assert abs(w.sum() - 2.0) < 1e-13
    for deg in range(0, min(No, 10) + 1):
        p = tau ** deg
        dp = deg * tau ** max(deg - 1, 0) if deg > 0 else np.zeros(N)
        assert np.abs(D @ p - dp).max() < 1e-8 * max(1, No) ** 2
    for deg in range(0, min(No, 9) + 1):  # Clenshaw-Curtis is exact to degree No
        exact = 0.0 if deg % 2 else 2.0 / (deg + 1)
        assert abs(w @ tau ** deg - exact) < 1e-12


def test_defect_of_exact_polynomial_trajectory_is_zero():
    """si2d with x(t) = a + b t, u = b satisfies the dynamics exactly: defects vanish."""
    wl = W.reference_vgp("ocp")
    o = ob.Oracle(wl)
    N = 33
    tau, w, D = o.collocation(0, N)
    t = 8.0 * tau + 8.0
    x = wl.x.copy()
    for k in range(N):
        x[0, wl.ix(0, k, 0)] = 1.0 + 0.25 * t[k]
        x[0, wl.ix(0, k, 1)] = 2.0 + 0.125 * t[k]
        x[0, wl.iu(0, k, 0)] = 0.25
        x[0, wl.iu(0, k, 1)] = 0.125
    r = o.eval(x, want=("f", "g"))
    assert np.abs(r["g"][0][:2 * N]).max() < 1e-11
    # objective = integral of (u0^2+u1^2) over [0,16]
    assert abs(r["f"][0] - 16.0 * (0.25 ** 2 + 0.125 ** 2)) < 1e-12
    # events = [x(t0); x(tf)], last row = tf - t0
    g = r["g"][0]
    assert list(g[2 * N:2 * N + 4]) == [1.0, 2.0, 1.0 + 0.25 * 16.0, 2.0 + 0.125 * 16.0]
    assert g[-1] == 16.0


WORKLOADS = [
    lambda: W.reference_vgp("ocp", batch=2, jitter=0.02),
    lambda: W.reference_vgp("mip", batch=1),
    lambda: W.pm3d(batch=2, nnodes=12, ncyl=3),
    lambda: W.pm3d(batch=1, nnodes=12, ncyl=3, scaled=True, pattern_mode=W.MODEL_DEPS),
    lambda: W.fw6(batch=1, nnodes=15, ncyl=4, scaled=True),
    lambda: W.pm3d_multiphase(batch=1, nnodes=9, ncyl=2, scaled=True),
    lambda: W.reference_vgp("ocp", collocation=W.CHEBYSHEV, maximize=True),
]


@pytest.mark.parametrize("mk", WORKLOADS)
def test_fd_jacobian_approximates_exact(mk):
    """numerical mode ~ automatic mode (PSOPT derivatives, src/ePSOPT/ePSOPT.cpp:64)."""
    wl = mk()
    o = ob.Oracle(wl)
    je = o.eval(wl.x, want=("jac",), jac_mode=0)["jac"]
    jf = o.eval(wl.x, want=("jac",), jac_mode=1)["jac"]
    scale = np.abs(je).max()
    assert np.abs(je - jf).max() <= 2e-6 * scale


@pytest.mark.parametrize("mk", WORKLOADS)
def test_reference_style_and_tight_style_agree_bitwise(mk):
    wl = mk()
    o = ob.Oracle(wl)
    a = o.eval(wl.x, want=("f", "g", "jac"), jac_mode=1, style=0)
    b = o.eval(wl.x, want=("f", "g", "jac"), jac_mode=1, style=1)
    for k in ("f", "g", "jac"):
        assert np.array_equal(a[k], b[k]), k


def test_exact_jacobian_matches_dymos_example_partials():
    """d(ellipse)/d(x,y) as hand-written in src/Examples/Dymos/etol_dymos_example1.cpp:239-240
    (without that file's exp() wrapping) and d(circle)/d(x,y) = -2 dx, -2 dy (:296-297)."""
    wl = W.reference_vgp("ocp")
    o = ob.Oracle(wl)
    irow, jcol, _ = o.structure()
    N, p0 = 33, 2 * 33 + 4
    jac = o.eval(wl.x, want=("jac",), jac_mode=0)["jac"][0]
    x = wl.x[0]
    recs = []
    for poly in W.REF_BORDERS:
        n = len(poly)
        for e in range(n):
            recs.append(ob.edge_geometry(poly[e], poly[(e + 1) % n]))
    for k in (0, 7, 16, 32):
        xk, yk = x[wl.ix(0, k, 0)], x[wl.ix(0, k, 1)]
        for q, g in enumerate(recs):
            xc, yc, radsq, tt = g[:4]
            asq, bsq = radsq, .2 * radsq
            dx, dy = xk - xc, yk - yc
            delx = np.cos(tt) * dx - np.sin(tt) * dy
            dely = np.sin(tt) * dx + np.cos(tt) * dy
            want = [-2. * (bsq * delx * np.cos(tt) + asq * dely * np.sin(tt)),
                    -2. * (-bsq * delx * np.sin(tt) + asq * dely * np.cos(tt))]
            for j in range(2):
                e = np.nonzero((irow == p0 + k * 11 + q) & (jcol == wl.ix(0, k, j)))[0]
                assert e.size == 1
                assert abs(jac[e[0]] - want[j]) <= 1e-12 * max(1.0, abs(want[j]))


def test_pattern_covers_every_nonzero_derivative():
    """dense finite differences of g: no entry outside the declared pattern may be non-zero."""
    for mk in (lambda: W.reference_vgp("mip"), lambda: W.fw6(batch=1, nnodes=7, ncyl=2, pattern_mode=W.MODEL_DEPS),
               lambda: W.pm3d_multiphase(batch=1, nnodes=5, ncyl=2, pattern_mode=W.MODEL_DEPS)):
        wl = mk()
        o = ob.Oracle(wl)
        irow, jcol, _ = o.structure()
        inpat = np.zeros((o.ncons, o.nvars), dtype=bool)
        inpat[irow, jcol] = True
        x0 = wl.x[0]
        for c in range(o.nvars):
            h = 1e-6 * (1 + abs(x0[c]))
            xp, xm = x0.copy(), x0.copy()
            xp[c] += h
            xm[c] -= h
            d = o.eval(xp, want=("g",), style=1)["g"][0] - o.eval(xm, want=("g",), style=1)["g"][0]
            assert not np.any((d != 0.0) & ~inpat[:, c]), f"column {c}"


def test_scaling_semantics():
    """solver sees z~ = z*sz, g~ = g*sg, f~ = f*sf."""
    wl = W.pm3d(batch=1, nnodes=10, ncyl=2)
    o = ob.Oracle(wl)
    base = o.eval(wl.x, want=("f", "g", "jac", "grad"), jac_mode=0)
    rng = np.random.default_rng(1)
    sz = 2.0 ** rng.integers(-3, 4, o.nvars).astype(float)  # powers of two: exact rescaling
    sg = 2.0 ** rng.integers(-3, 4, o.ncons).astype(float)
    o.set_scaling(sz, sg, 4.0)
    sc = o.eval(wl.x * sz, want=("f", "g", "jac", "grad"), jac_mode=0)
    irow, jcol, _ = o.structure()
    assert np.array_equal(sc["f"], 4.0 * base["f"])
    assert np.array_equal(sc["g"], base["g"] * sg)
    assert np.array_equal(sc["jac"], base["jac"] * sg[irow] / sz[jcol])
    assert np.array_equal(sc["grad"], 4.0 * base["grad"] / sz)


def test_gradient_matches_finite_differences():
    wl = W.fw6(batch=1, nnodes=9, ncyl=2)
    o = ob.Oracle(wl)
    grad = o.eval(wl.x, want=("grad",))["grad"][0]
    x0 = wl.x[0]
    num = np.zeros_like(grad)
    for c in range(o.nvars):
        h = 1e-6 * (1 + abs(x0[c]))
        xp, xm = x0.copy(), x0.copy()
        xp[c] += h
        xm[c] -= h
        num[c] = (o.eval(xp, want=("f",))["f"][0] - o.eval(xm, want=("f",))["f"][0]) / (2 * h)
    assert np.abs(num - grad).max() <= 1e-6 * max(1.0, np.abs(grad).max())
