"""Pins the ETOL-side rows of the hot path (a1 dae, a2 integrand_cost, a3 endpoint_cost, a4 events,
a6 setup / addBounds, a7-a10 the example's objective, dynamics, obsConstraint and saaConstraint) to the REFERENCE'S OWN
CODE: /root/reference/src/ePSOPT/ePSOPT.cpp and src/Examples/PSOPT/etol_psopt_example1.cpp compiled unmodified
against oracle/refstub/psopt.h (oracle/_ref/libetol_ref.so, `make -C oracle ref`).

Two layers:
  * where oracle/_ref exists (this container; it also travels to the GPU box as a built file), the reference
    callbacks are executed at the parity tests' node inputs and compared with the oracle's values;
  * everywhere, the oracle is compared with tests/golden/c0_ref_pernode.npz, outputs of those same reference
    callbacks committed by tests/golden/make_golden.py.
What stays unpinned is the PSOPT side (D, quadrature, layout, scaling, colouring, FD step): PSOPT is not in the
reference tree. The moving-zone rows go through PSOPT's linear_interpolation, which the stub restates with the rule of
ETOL's own TrajectoryOptimizer::linear_interpolation."""
import os

import numpy as np
import pytest

import oracle_binding as ob
import ref_binding as rb
from etol_b200 import workloads as W

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c0_ref_pernode.npz")
VAL = 1e-12  # north_star tolerance on values


def _xml(tmp_path):
    if os.path.exists(rb.REF_XML):
        return rb.REF_XML  # the shipped file itself
    import plugin_binding as pb
    return pb.write_reference_xml(str(tmp_path / "ocp_2d_ex1.xml"), "ocp")


def _close(a, b, tol, scale=None):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    s = np.maximum(np.abs(a), np.abs(b)) if scale is None else scale
    return bool(np.all(np.abs(a - b) <= tol * np.maximum(s, 1e-300)))


def _oracle_pieces(wl, o, x, b=0):
    """per-node pieces of the oracle's evaluation of instance b (all-ones scaling)"""
    N, ns, npath = wl.nnodes[0], wl.ns, wl.npath[0]
    r = o.eval(x, want=("f", "g", "jac"), jac_mode=W.JAC_EXACT)
    g = r["g"][b]
    ne = 2 * ns
    return dict(f=r["f"][b], defects=g[:ns * N].reshape(N, ns), events=g[ns * N:ns * N + ne],
                path=g[ns * N + ne:ns * N + ne + npath * N].reshape(N, npath), duration=g[-1], jac=r["jac"][b])


def _compare(wl, x, ref_nodes, b=0):
    """ref_nodes: dict with the reference callbacks' outputs at the node inputs of instance b"""
    o = ob.Oracle(wl)
    N, ns, nc, npath = wl.nnodes[0], wl.ns, wl.nc, wl.npath[0]
    X, U, t0, tf = rb.node_inputs(wl, x, b)
    tau, w, D = o.collocation(0, N)
    h = 0.5 * (tf - t0)
    pc = _oracle_pieces(wl, o, x, b)
    # a1 + a8 (dae: state derivatives) through the defect rows  zeta = D X - h f
    zeta_ref = D @ X - h * ref_nodes["f"]
    scale = np.abs(D) @ np.abs(X) + abs(h) * np.abs(ref_nodes["f"])
    assert _close(pc["defects"], zeta_ref, 4 * VAL, scale), "defect rows vs reference dae"
    # a1 + a9 + a10 (dae: path rows = obsConstraint edges then saaConstraint tracks), value by value
    assert pc["path"].shape == ref_nodes["path"].shape
    assert _close(pc["path"], ref_nodes["path"], VAL), "path rows vs reference obs / saa lambdas"
    # a4 events, exact
    assert np.array_equal(pc["events"], ref_nodes["events"]), "event rows vs reference events()"
    # a2 + a7 + a13 objective = h * sum_k w_k L_k with the reference's integrand; a3 endpoint cost = 0
    assert ref_nodes["endpoint"] == 0.0
    f_ref = h * float(np.dot(w, ref_nodes["L"]))
    assert abs(pc["f"] - f_ref) <= 4 * VAL * max(abs(f_ref), 1e-300), "objective vs reference integrand_cost"
    # exact Jacobian (the reference's default, derivatives = "automatic"): path-row and dynamics partials
    irow, jcol, _ = o.structure()
    ent = {(int(r), int(c)): v for r, c, v in zip(irow, jcol, pc["jac"])}
    ne = 2 * ns
    worst = 0.0
    for k in range(N):
        a, bb = 0.5 * (1.0 - tau[k]), 0.5 * (1.0 + tau[k])  # d t_k / d t0, d t_k / d tf
        for q in range(npath):
            r = ns * N + ne + k * npath + q
            dp = ref_nodes["dpath"][k, q]
            for j in range(ns):
                got = ent.get((r, wl.ix(0, k, j)), 0.0)
                worst = max(worst, abs(got - dp[j]) / max(abs(dp[j]), 1.0))
            dt = dp[ns + nc]
            if dt != 0.0 or (r, wl.it0(0)) in ent:
                worst = max(worst, abs(ent.get((r, wl.it0(0)), 0.0) - dt * a) / max(abs(dt * a), 1.0))
                worst = max(worst, abs(ent.get((r, wl.itf(0)), 0.0) - dt * bb) / max(abs(dt * bb), 1.0))
        for i in range(ns):
            r = k * ns + i
            df = ref_nodes["df"][k, i]
            for j in range(nc):  # d zeta_ki / d u_kj = -h df_i/du_j
                got = ent.get((r, wl.iu(0, k, j)), 0.0)
                worst = max(worst, abs(got - (-h * df[ns + j])) / max(abs(h * df[ns + j]), 1.0))
            for j in range(ns):  # d zeta_ki / d x_kj = D_kk delta_ij - h df_i/dx_j
                got = ent.get((r, wl.ix(0, k, j)), 0.0)
                want = (D[k, k] if i == j else 0.0) - h * df[j]
                worst = max(worst, abs(got - want) / max(abs(want), 1.0))
    assert worst <= 1e-12, f"exact Jacobian vs the reference callbacks' tangents: {worst:.3e}"


def _run_reference(ref, wl, x, b=0):
    N = wl.nnodes[0]
    X, U, t0, tf = rb.node_inputs(wl, x, b)
    o = ob.Oracle(wl)
    tau, _, _ = o.collocation(0, N)
    h, m = 0.5 * (tf - t0), 0.5 * (tf + t0)
    t = h * tau + m  # the oracle's node times, same expression
    out = dict(f=[], path=[], df=[], dpath=[], L=[], t=t)
    for k in range(N):
        f, p, df, dp = ref.dae(X[k], U[k], t[k])
        L, _ = ref.integrand_cost(X[k], U[k], t[k])
        out["f"].append(f), out["path"].append(p), out["df"].append(df), out["dpath"].append(dp), out["L"].append(L)
    out = {k: np.array(v) for k, v in out.items()}
    out["events"] = ref.events(X[0], X[-1], t0, tf)
    out["endpoint"] = ref.endpoint_cost(X[0], X[-1], t0, tf)
    return out


needs_ref = pytest.mark.skipif(not rb.available(), reason="oracle/_ref/libetol_ref.so not built (no /root/reference here)")


@needs_ref
def test_reference_setup_dimensions_bounds_and_options(tmp_path):
    """ePSOPT::setup / addBounds of the reference itself against the workload the parity tests use (row a6)"""
    ref = rb.Reference(_xml(tmp_path))
    wl = W.reference_vgp("ocp")
    N = wl.nnodes[0]
    assert (ref.ns, ref.nc, ref.ne, ref.npath, ref.nodes) == (wl.ns, wl.nc, 2 * wl.ns, wl.npath[0], N) == (2, 2, 4, 11, 33)
    b = ref.bounds()
    ns, ne, npath = wl.ns, 2 * wl.ns, wl.npath[0]
    gl, gu = wl.gl[0], wl.gu[0]
    assert np.array_equal(b["events"][0], gl[ns * N:ns * N + ne]) and np.array_equal(b["events"][1], gu[ns * N:ns * N + ne])
    for k in range(N):
        sl = slice(ns * N + ne + k * npath, ns * N + ne + (k + 1) * npath)
        assert np.array_equal(b["path"][0], gl[sl]) and np.array_equal(b["path"][1], gu[sl])
        for i in range(ns):
            assert (wl.zl[wl.ix(0, k, i)], wl.zu[wl.ix(0, k, i)]) == (b["states"][0][i], b["states"][1][i])
        for j in range(wl.nc):
            assert (wl.zl[wl.iu(0, k, j)], wl.zu[wl.iu(0, k, j)]) == (b["controls"][0][j], b["controls"][1][j])
    assert (wl.zl[wl.it0(0)], wl.zu[wl.it0(0)]) == tuple(b["t0"]) == (0.0, 0.0)
    assert (wl.zl[wl.itf(0)], wl.zu[wl.itf(0)]) == tuple(b["tf"]) == (16.0, 16.0)
    strings, nums = ref.algorithm()
    assert strings == ["IPOPT", "automatic", "automatic", "exact", "Legendre", "automatic"]
    assert list(nums) == [200.0, 1e-6, 10.0, 1e-4, 0.0]
    assert np.array_equal(ref.guess_time(), np.linspace(0.0, 16.0, N))
    ref.close()


@needs_ref
def test_plugin_bounds_match_reference_addbounds(tmp_path):
    """eCUDA::buildBounds (the product's plugin) against ePSOPT::addBounds run from the same XML"""
    import plugin_binding as pb
    xml = _xml(tmp_path)
    ref = rb.Reference(xml)
    p = pb.Plugin().load(xml, model=W.SI2D, scaling="none")
    wl = W.reference_vgp("ocp")
    bd, rbd = p.bounds(), ref.bounds()
    N, ns, ne, npath = wl.nnodes[0], wl.ns, 2 * wl.ns, wl.npath[0]
    assert np.array_equal(bd["gl"][ns * N:ns * N + ne], rbd["events"][0]) and np.array_equal(bd["gu"][ns * N:ns * N + ne], rbd["events"][1])
    assert np.array_equal(bd["gl"][ns * N + ne:ns * N + ne + npath], rbd["path"][0])
    assert np.array_equal(bd["gu"][ns * N + ne:ns * N + ne + npath], rbd["path"][1])
    assert np.all(bd["gl"][:ns * N] == 0.0) and np.all(bd["gu"][:ns * N] == 0.0)
    for k in (0, N // 2, N - 1):
        for i in range(ns):
            assert (bd["zl"][wl.ix(0, k, i)], bd["zu"][wl.ix(0, k, i)]) == (rbd["states"][0][i], rbd["states"][1][i])
        for j in range(wl.nc):
            assert (bd["zl"][wl.iu(0, k, j)], bd["zu"][wl.iu(0, k, j)]) == (rbd["controls"][0][j], rbd["controls"][1][j])
    assert (bd["zl"][wl.itf(0)], bd["zu"][wl.itf(0)]) == tuple(rbd["tf"])
    p.close()
    ref.close()


@needs_ref
@pytest.mark.parametrize("seed", [0xE701, 7, 20261018])
def test_oracle_vs_reference_callbacks(tmp_path, seed):
    """rows a1-a4, a7-a10: the reference's dae / integrand_cost / events / endpoint_cost and its obs / saa lambdas,
    executed at every node of seeded decision vectors, against the oracle's values and exact partials"""
    ref = rb.Reference(_xml(tmp_path))
    wl = W.reference_vgp("ocp", seed=seed)
    _compare(wl, wl.x, _run_reference(ref, wl, wl.x))
    ref.close()


@needs_ref
def test_reference_maximize_negates_the_integrand(tmp_path):
    ref_min, ref_max = rb.Reference(_xml(tmp_path), False), rb.Reference(_xml(tmp_path), True)
    x, u = np.array([1.5, 2.5]), np.array([0.25, -0.125])
    assert ref_max.integrand_cost(x, u, 3.0)[0] == -ref_min.integrand_cost(x, u, 3.0)[0] == -(0.25 ** 2 + 0.125 ** 2)
    wl = W.reference_vgp("ocp", maximize=True)
    out = _run_reference(ref_max, wl, wl.x)
    _compare(wl, wl.x, out)
    ref_min.close(), ref_max.close()


def test_oracle_matches_committed_reference_vectors():
    """the same comparison against tests/golden/c0_ref_pernode.npz (generated from oracle/_ref by make_golden.py):
    runs where neither /root/reference nor oracle/_ref exists"""
    gold = np.load(GOLD)
    wl = W.reference_vgp("ocp")
    assert np.array_equal(gold["x"], wl.x), "the committed decision vector is the seeded one of the parity tests"
    ref_nodes = {k: gold[k] for k in ("f", "path", "df", "dpath", "L", "events")}
    ref_nodes["endpoint"] = float(gold["endpoint"])
    _compare(wl, wl.x, ref_nodes)
    N, ns, ne, npath = wl.nnodes[0], wl.ns, 2 * wl.ns, wl.npath[0]
    assert np.array_equal(gold["bounds_events"][0], wl.gl[0][ns * N:ns * N + ne])
    assert np.array_equal(gold["bounds_events"][1], wl.gu[0][ns * N:ns * N + ne])
    assert np.array_equal(gold["bounds_path"][0], wl.gl[0][ns * N + ne:ns * N + ne + npath])
    assert np.array_equal(gold["bounds_path"][1], wl.gu[0][ns * N + ne:ns * N + ne + npath])
