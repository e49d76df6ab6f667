"""Parity tests proper: the CUDA path, called through the C ABI (include/ecuda.h), against the CPU
oracle on the same seeded inputs. Bars (BASELINE.json north_star): sparsity pattern and colouring
bit-exact; f and g within 1e-12 relative; Jacobian entries within 1e-9 relative under the same
perturbation step. FD mode additionally asserts identical bits (same IEEE operation sequence)."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_binding as ob
from conftest import TOL_JAC, TOL_VALUE, rel_err
from etol_b200 import capi, workloads as W

pytestmark = pytest.mark.gpu

CASES = {
    # C0: the reference VGP (ocp_2d_ex1.xml shape) and the mip_2d_ex1.xml variant the Singularity app runs
    "C0-ocp": lambda: W.reference_vgp("ocp", batch=5, jitter=0.02),
    "C0-mip": lambda: W.reference_vgp("mip", batch=3, jitter=0.01),
    "C0-ocp-cheb-max": lambda: W.reference_vgp("ocp", collocation=W.CHEBYSHEV, maximize=True),
    "C0-ocp-deps-base1": lambda: W.reference_vgp("ocp", pattern_mode=W.MODEL_DEPS, index_base=1),
    # C1/C2: 3-D point mass, 8 cylinders, 40 LGL nodes
    "C1-pm3d": lambda: W.pm3d(batch=1),
    "C2-pm3d-64": lambda: W.pm3d(batch=64),
    "C2-pm3d-scaled-deps": lambda: W.pm3d(batch=7, scaled=True, pattern_mode=W.MODEL_DEPS),
    # C3: fixed wing, 200 nodes, 64 cylinders
    "C3-fw6": lambda: W.fw6(batch=2),
    "C3-fw6-small-scaled": lambda: W.fw6(batch=3, nnodes=21, ncyl=5, scaled=True),
    # C4: three linked phases
    "C4-multiphase": lambda: W.pm3d_multiphase(batch=6, scaled=True),
    "C4-multiphase-ragged": lambda: W.pm3d_multiphase(batch=2, nphases=4, nnodes=7, ncyl=1),
    # edge cases: minimum node count, no obstacles, node counts off the block size
    "pm3d-N2": lambda: W.pm3d(batch=2, nnodes=2, ncyl=1),
    "pm3d-no-obstacles": lambda: W.pm3d(batch=2, nnodes=9, ncyl=0),
    "pm3d-N65": lambda: W.pm3d(batch=2, nnodes=65, ncyl=3),
    # user models: dynamics and cost recorded from callbacks, kernels compiled at run time (NVRTC)
    "user-pm3d": lambda: W.pm3d_user(batch=9),
    "user-unicycle-tracks": lambda: W.unicycle(batch=6, ntracks=2),
    "user-unicycle-cheb-deps": lambda: W.unicycle(batch=3, nnodes=40, collocation=W.CHEBYSHEV, pattern_mode=W.MODEL_DEPS,
                                                 scaled=True),
    "user-dragmass": lambda: W.dragmass(batch=5, ntracks=1, scaled=True),
    "user-dragmass-N12-generic": lambda: W.dragmass(batch=3, nnodes=12, ncyl=2),
    "user-dragmass-N70-generic": lambda: W.dragmass(batch=2, nnodes=70, ncyl=3, maximize=True),
    # user models whose dynamics and running cost read t (VERDICT r1 missing item 3)
    "user-gust": lambda: W.gust(batch=5, ntracks=1, scaled=True),
    "user-gust-cheb-deps": lambda: W.gust(batch=3, nnodes=40, collocation=W.CHEBYSHEV, pattern_mode=W.MODEL_DEPS),
    "user-gust-N70-generic": lambda: W.gust(batch=2, nnodes=70, ncyl=3, maximize=True),
    # traced path constraints: rows that are none of the built-in zone rows (VERDICT r1 missing item 2)
    "user-zone": lambda: W.zone(batch=5, ntracks=2, scaled=True),
    "user-zone-gust-cheb": lambda: W.zone(batch=3, timedep=True, nnodes=40, ncyl=0, collocation=W.CHEBYSHEV),
    "user-zone-N70-generic": lambda: W.zone(batch=2, nnodes=70, ncyl=3, pattern_mode=W.MODEL_DEPS),
}


@pytest.fixture(scope="module")
def evaluators():
    cache = {}
    yield cache
    for ev, _, _ in cache.values():
        ev.close()


def _get(evaluators, name):
    if name not in evaluators:
        wl = CASES[name]()
        evaluators[name] = (capi.Evaluator(wl, device=0), ob.Oracle(wl), wl)
    return evaluators[name]


@pytest.mark.parametrize("name", sorted(CASES))
def test_structure_bit_exact(evaluators, name):
    ev, orc, wl = _get(evaluators, name)
    for a, b in zip(ev.structure(), orc.structure()):
        assert np.array_equal(a, b)
    assert (ev.nvars, ev.ncons, ev.nnz, ev.dims.ngroups) == (orc.nvars, orc.ncons, orc.nnz, orc.ngroups)
    for p, N in enumerate(wl.nnodes):
        for a, b in zip(ev.collocation(p), orc.collocation(p, N)):
            assert np.array_equal(a, b)


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("mode", [W.JAC_FD, W.JAC_EXACT])
def test_values_match_oracle(evaluators, name, mode):
    ev, orc, wl = _get(evaluators, name)
    style = 1 if name == "C3-fw6" else 0  # tight oracle for the 200-node case (bit-identical to style 0)
    ref = orc.eval(wl.x, want=("f", "g", "jac", "grad"), jac_mode=mode, style=style, nthreads=ob.max_threads())
    got = ev.eval_host(wl.x, want=("f", "g", "jac", "grad"), jac_mode=mode)
    assert rel_err(got["f"], ref["f"]) <= TOL_VALUE
    assert rel_err(got["g"], ref["g"]) <= TOL_VALUE
    assert rel_err(got["grad"], ref["grad"]) <= TOL_VALUE
    assert rel_err(got["jac"], ref["jac"]) <= TOL_JAC
    if mode == W.JAC_FD:
        assert np.array_equal(got["f"], ref["f"]) and np.array_equal(got["g"], ref["g"])
        assert np.array_equal(got["jac"], ref["jac"]), "FD Jacobian is not bit-identical to the oracle"


MESH_CASES = ["C0-ocp", "C0-ocp-cheb-max", "C2-pm3d-scaled-deps", "C3-fw6", "C3-fw6-small-scaled", "C4-multiphase",
              "C4-multiphase-ragged", "pm3d-N2", "pm3d-N65", "user-unicycle-tracks", "user-dragmass",
              "user-dragmass-N70-generic", "user-gust", "user-gust-N70-generic"]


@pytest.mark.parametrize("name", MESH_CASES)
def test_ode_error_matches_oracle(evaluators, name):
    """relative local discretisation error per mesh interval (mesh-refinement input)"""
    ev, orc, wl = _get(evaluators, name)
    ref = ob.ode_error(orc, wl, wl.x)
    got = ev.ode_error_host(wl.x)
    assert got.shape == ref.shape and np.isfinite(got).all() and (got >= 0).all()
    assert rel_err(got, ref) <= TOL_VALUE


@pytest.mark.parametrize("name", MESH_CASES)
def test_resample_matches_oracle(evaluators, name):
    """decision vectors interpolated onto a finer and a coarser mesh, unscaled and re-scaled"""
    ev, orc, wl = _get(evaluators, name)
    for nn in ([n + 7 for n in wl.nnodes], [max(2, n // 2) for n in wl.nnodes]):
        nv = sum((wl.ns + wl.nc) * n + 2 for n in nn)
        sz = np.linspace(0.5, 2.0, nv)
        for s in (None, sz):
            ref = ob.resample(orc, wl, wl.x, nn, sz_new=s)
            got = ev.resample_host(wl.x, nn, sz_new=s)
            assert got.shape == ref.shape
            assert rel_err(got, ref) <= TOL_VALUE


HESS_CASES = ["C0-ocp", "C0-mip", "C0-ocp-cheb-max", "C0-ocp-deps-base1", "C2-pm3d-64", "C2-pm3d-scaled-deps", "C3-fw6",
              "C3-fw6-small-scaled", "C4-multiphase", "C4-multiphase-ragged", "pm3d-N2", "pm3d-no-obstacles",
              "user-pm3d", "user-unicycle-tracks", "user-unicycle-cheb-deps", "user-dragmass", "user-dragmass-N70-generic",
              "user-gust", "user-gust-cheb-deps", "user-zone", "user-zone-gust-cheb", "user-zone-N70-generic"]


@pytest.mark.parametrize("name", HESS_CASES)
def test_hessian_matches_oracle(evaluators, name):
    """Hessian of the Lagrangian (lower triangle) against the oracle's second-order forward mode"""
    ev, orc, wl = _get(evaluators, name)
    rng = np.random.default_rng(7)
    lam = rng.normal(size=(wl.batch, ev.ncons))
    sigma = rng.uniform(0.25, 2.0, size=wl.batch)
    n = C.c_int32(0)
    assert capi.lib().ecuda_get_hess_structure(ev.h, C.byref(n), None, None) == 0
    irow, jcol = np.zeros(n.value, np.int32), np.zeros(n.value, np.int32)
    assert capi.lib().ecuda_get_hess_structure(ev.h, None, irow.ctypes.data_as(capi._ip), jcol.ctypes.data_as(capi._ip)) == 0
    oi, oj = ob.hess_structure(orc)
    assert np.array_equal(irow, oi) and np.array_equal(jcol, oj)
    ref = ob.eval_hess(orc, wl.x, sigma, lam)
    got = ev.hess_host(wl.x, sigma, lam)
    assert got.shape == ref.shape
    assert rel_err(got, ref) <= TOL_JAC
    # one objective factor for the whole batch (the IPOPT-shaped call uses this form)
    if wl.batch == 1:
        vals = np.zeros(n.value)
        x0, l0 = np.ascontiguousarray(wl.x[0]), np.ascontiguousarray(lam[0])
        rc = capi.lib().ecuda_ipopt_eval_h(ev.h, ev.nvars, x0.ctypes.data_as(capi._dp), 1, float(sigma[0]), ev.ncons,
                                           l0.ctypes.data_as(capi._dp), 1, n.value, None, None, vals.ctypes.data_as(capi._dp))
        assert rc == 0 and np.array_equal(vals, got[0])


def test_hessian_error_and_resample_with_device_pointers(evaluators):
    """the same three entry points with device buffers on a caller's stream (no staging copies)"""
    import torch
    ev, orc, wl = _get(evaluators, "C2-pm3d-scaled-deps")
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(3)
    lam = rng.normal(size=(wl.batch, ev.ncons))
    sigma = rng.uniform(0.5, 1.5, size=wl.batch)
    stream = torch.cuda.Stream(device=dev)
    st = stream.cuda_stream
    with torch.cuda.stream(stream):
        x = torch.from_numpy(wl.x).to(dev)
        lam_d, sig_d = torch.from_numpy(lam).to(dev), torch.from_numpy(sigma).to(dev)
        n = C.c_int32(0)
        assert capi.lib().ecuda_get_hess_structure(ev.h, C.byref(n), None, None) == 0
        hv = torch.full((wl.batch, n.value), float("nan"), dtype=torch.float64, device=dev)
        assert capi.lib().ecuda_eval_hess(ev.h, x.data_ptr(), sig_d.data_ptr(), 0.0, lam_d.data_ptr(), hv.data_ptr(),
                                          capi.MEM_DEVICE, st) == 0
        nint = sum(k - 1 for k in wl.nnodes)
        err = torch.full((wl.batch, nint), float("nan"), dtype=torch.float64, device=dev)
        assert capi.lib().ecuda_ode_error(ev.h, x.data_ptr(), err.data_ptr(), capi.MEM_DEVICE, st) == 0
        nn = np.array([n_ + 5 for n_ in wl.nnodes], dtype=np.int32)
        nv = sum((wl.ns + wl.nc) * int(k) + 2 for k in nn)
        xn = torch.full((wl.batch, nv), float("nan"), dtype=torch.float64, device=dev)
        assert capi.lib().ecuda_resample(ev.h, x.data_ptr(), nn.ctypes.data_as(capi._ip), None, xn.data_ptr(),
                                         capi.MEM_DEVICE, st) == 0
    stream.synchronize()
    assert np.array_equal(hv.cpu().numpy(), ev.hess_host(wl.x, sigma, lam))
    assert np.array_equal(err.cpu().numpy(), ev.ode_error_host(wl.x))
    assert np.array_equal(xn.cpu().numpy(), ev.resample_host(wl.x, nn))


@pytest.mark.parametrize("uniform_bounds", [True, False])
@pytest.mark.parametrize("mode", [W.JAC_FD, W.JAC_EXACT])
def test_fused_summary_equals_separate_summary(uniform_bounds, mode):
    """the {f, max bound violation} rows the evaluation kernel's epilogue writes (one rank: into its own buffer)
    are the bits of the stand-alone summary kernel -- with the compact form of the bounds (every instance has
    defect bounds 0 and the same path-row bounds) and with dense per-instance bounds"""
    import torch
    wl = W.pm3d(batch=37)
    if not uniform_bounds:  # one instance with its own path-row bound, one with a non-zero defect bound
        p0 = wl.ns * wl.nnodes[0] + 2 * wl.ns
        wl.gl[5, p0 + 3] = -2.5
        wl.gu[9, 4] = 0.125
    ev = capi.Evaluator(wl, device=0)
    dev = torch.device("cuda:0")
    stream = torch.cuda.Stream(device=dev)
    st = stream.cuda_stream
    with torch.cuda.stream(stream):
        x = torch.from_numpy(wl.x).to(dev)
        f = torch.empty(wl.batch, dtype=torch.float64, device=dev)
        g = torch.empty((wl.batch, ev.ncons), dtype=torch.float64, device=dev)
        jac = torch.empty((wl.batch, ev.nnz), dtype=torch.float64, device=dev)
        fused = torch.full((wl.batch, 2), float("nan"), dtype=torch.float64, device=dev)
        sep = torch.full((wl.batch, 2), float("nan"), dtype=torch.float64, device=dev)
        ev.eval_allgather_ptr(x.data_ptr(), f.data_ptr(), g.data_ptr(), jac.data_ptr(), mode, [fused.data_ptr()], 0, st)
        ev.summarize_ptr(f.data_ptr(), g.data_ptr(), sep.data_ptr(), st)
    stream.synchronize()
    a, b = fused.cpu().numpy(), sep.cpu().numpy()
    assert np.array_equal(a, b)
    gv = g.cpu().numpy()
    want = np.maximum(wl.gl - gv, gv - wl.gu).max(axis=1)   # the definition, on the host
    assert np.array_equal(a[:, 1], want) and np.array_equal(a[:, 0], f.cpu().numpy())
    ev.close()


def test_resample_round_trip_on_device(evaluators):
    """up-sampling does not change the interpolating polynomial: 40 -> 61 -> 40 nodes returns the decision
    vectors (a size-independent property of the kernel), and the finer mesh sees an error profile of the
    same size"""
    ev, _, wl = _get(evaluators, "C2-pm3d-64")
    fine = W.pm3d(batch=wl.batch, nnodes=61)
    ev2 = capi.Evaluator(fine, device=0)
    up = ev.resample_host(wl.x, [61])
    back = ev2.resample_host(up, [40])
    assert np.abs(back - wl.x).max() <= 1e-9 * np.abs(wl.x).max()
    e1, e2 = ev.ode_error_host(wl.x), ev2.ode_error_host(up)
    assert e1.shape == (wl.batch, 39) and e2.shape == (wl.batch, 60)
    ratio = e2.sum(axis=1) / e1.sum(axis=1)   # sum of per-interval maxima: comparable, not identical
    assert (ratio > 0.8).all() and (ratio < 1.5).all()
    ev2.close()


def test_user_model_equals_builtin_bitwise(evaluators):
    """pm3d written as callbacks and compiled at run time gives the bits of the built-in pm3d kernels"""
    ev_u, _, wl_u = _get(evaluators, "user-pm3d")
    wl_b = W.pm3d(batch=wl_u.batch)
    ev_b = capi.Evaluator(wl_b, device=0)
    assert np.array_equal(wl_u.x, wl_b.x)
    for mode in (W.JAC_FD, W.JAC_EXACT):
        a = ev_u.eval_host(wl_u.x, want=("f", "g", "jac", "grad"), jac_mode=mode)
        b = ev_b.eval_host(wl_b.x, want=("f", "g", "jac", "grad"), jac_mode=mode)
        for k in ("f", "g", "jac", "grad"):
            assert np.array_equal(a[k], b[k]), (mode, k)
    ev_b.close()


def test_golden_fixture_c0(evaluators):
    """committed oracle outputs for the reference VGP at a committed decision vector"""
    import os
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "c0_ocp_golden.npz"))
    wl = W.reference_vgp("ocp")
    assert np.array_equal(wl.x, gold["x"]), "workload generator drifted from the committed fixture"
    ev = capi.Evaluator(wl)
    fd = ev.eval_host(wl.x, jac_mode=W.JAC_FD)
    ex = ev.eval_host(wl.x, want=("jac",), jac_mode=W.JAC_EXACT)
    assert rel_err(fd["f"], gold["f"]) <= TOL_VALUE and rel_err(fd["g"], gold["g"]) <= TOL_VALUE
    assert rel_err(fd["jac"], gold["jac_fd"]) <= TOL_JAC and rel_err(ex["jac"], gold["jac_exact"]) <= TOL_JAC
    irow, jcol, grp = ev.structure()
    assert np.array_equal(irow, gold["irow"]) and np.array_equal(jcol, gold["jcol"])
    assert np.array_equal(grp, gold["group_of_col"])
    ev.close()


def test_extended_golden_fixture(evaluators):
    """committed oracle outputs: Hessian, discretisation error, resampling on C0; a runtime-compiled user model"""
    import os
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "c0_ext_golden.npz"))
    wl = W.reference_vgp("ocp")
    ev = capi.Evaluator(wl)
    assert rel_err(ev.hess_host(wl.x, gold["sigma"], gold["lam"]), gold["hess"]) <= TOL_JAC
    assert rel_err(ev.ode_error_host(wl.x), gold["ode_error"]) <= TOL_VALUE
    assert rel_err(ev.resample_host(wl.x, [41]), gold["resample_41"]) <= TOL_VALUE
    ev.close()
    uw = W.unicycle(batch=1, nnodes=17, ncyl=2, ntracks=1)
    assert np.array_equal(uw.x, gold["user_x"])
    uev = capi.Evaluator(uw)
    fd = uev.eval_host(uw.x, jac_mode=W.JAC_FD)
    ex = uev.eval_host(uw.x, want=("jac",), jac_mode=W.JAC_EXACT)
    assert rel_err(fd["f"], gold["user_f"]) <= TOL_VALUE and rel_err(fd["g"], gold["user_g"]) <= TOL_VALUE
    assert np.array_equal(fd["jac"], gold["user_jac_fd"]) and rel_err(ex["jac"], gold["user_jac_exact"]) <= TOL_JAC
    uev.close()


def test_device_pointer_path_and_partial_outputs(evaluators):
    import torch
    ev, orc, wl = _get(evaluators, "C2-pm3d-64")
    dev = torch.device("cuda:0")
    x = torch.from_numpy(wl.x).to(dev)
    f = torch.full((wl.batch,), float("nan"), dtype=torch.float64, device=dev)
    g = torch.full((wl.batch, ev.ncons), float("nan"), dtype=torch.float64, device=dev)
    jac = torch.full((wl.batch, ev.nnz), float("nan"), dtype=torch.float64, device=dev)
    stream = torch.cuda.Stream(device=dev)
    st = stream.cuda_stream
    stream.wait_stream(torch.cuda.current_stream())
    ev.eval_ptr(x.data_ptr(), f.data_ptr(), g.data_ptr(), jac.data_ptr(), W.JAC_FD, capi.MEM_DEVICE, st)
    stream.synchronize()
    ref = orc.eval(wl.x, jac_mode=W.JAC_FD, nthreads=ob.max_threads())
    assert np.array_equal(f.cpu().numpy(), ref["f"])
    assert np.array_equal(g.cpu().numpy(), ref["g"])
    assert np.array_equal(jac.cpu().numpy(), ref["jac"])
    # g only: f and jac buffers must stay untouched
    g2 = torch.zeros_like(g)
    ev.eval_ptr(x.data_ptr(), None, g2.data_ptr(), None, W.JAC_FD, capi.MEM_DEVICE, st)
    torch.cuda.synchronize()
    assert torch.equal(g2, g)


@pytest.mark.parametrize("name", ["C0-ocp", "C2-pm3d-64", "C2-pm3d-scaled-deps", "C4-multiphase", "C3-fw6-small-scaled",
                                  "pm3d-N2", "user-dragmass"])
def test_compact_exact_jacobian_splices_to_the_full_one(evaluators, name):
    """ecuda_eval_compact returns only the per-instance triplets of the exact Jacobian (26 % of them at the benchmark
    shape); shared values + splice must reproduce ecuda_eval's full array bit for bit, through host buffers and
    through device pointers, and f / g must be the same as ecuda_eval's."""
    import torch
    ev, orc, wl = _get(evaluators, name)
    full = ev.eval_host(wl.x, want=("f", "g", "jac"), jac_mode=W.JAC_EXACT)
    idx, shared = ev.compact_structure()
    assert np.array_equal(idx, capi.host_compact_structure(wl))
    assert 0 < idx.size < ev.nnz or min(wl.nnodes) < 2
    out = ev.eval_compact_host(wl.x, idx.size)
    assert np.array_equal(out["f"], full["f"]) and np.array_equal(out["g"], full["g"])
    assert np.array_equal(out["jac_local"], full["jac"][:, idx])
    assert np.array_equal(ev.splice(shared, idx, out["jac_local"]), full["jac"])
    dev = torch.device("cuda:0")
    x = torch.from_numpy(wl.x).to(dev)
    f = torch.empty(wl.batch, dtype=torch.float64, device=dev)
    g = torch.empty((wl.batch, ev.ncons), dtype=torch.float64, device=dev)
    jl = torch.full((wl.batch, idx.size), float("nan"), dtype=torch.float64, device=dev)
    ev.eval_compact_ptr(x.data_ptr(), f.data_ptr(), g.data_ptr(), jl.data_ptr(), capi.MEM_DEVICE, None)
    ev.sync()
    assert np.array_equal(jl.cpu().numpy(), out["jac_local"]) and np.array_equal(g.cpu().numpy(), full["g"])


@pytest.mark.parametrize("chunks", ["1", "3", "8"])
def test_chunked_host_path_gives_the_same_bits(evaluators, chunks):
    """HOST-buffer calls run in instance chunks on three streams (eval_host); any chunk count -- more chunks than
    instances included -- must return what the device-pointer path returns, for every output and both modes."""
    import torch
    for name in ("C2-pm3d-scaled-deps", "C4-multiphase", "user-zone"):
        wl = CASES[name]()
        os.environ["ECUDA_HOST_CHUNKS"] = chunks  # read by ecuda_create
        try:
            ev = capi.Evaluator(wl, device=0)
        finally:
            os.environ.pop("ECUDA_HOST_CHUNKS", None)
        dev = torch.device("cuda:0")
        x = torch.from_numpy(wl.x).to(dev)
        for mode in (W.JAC_FD, W.JAC_EXACT):
            got = ev.eval_host(wl.x, want=("f", "g", "jac", "grad"), jac_mode=mode)
            f = torch.empty(wl.batch, dtype=torch.float64, device=dev)
            g = torch.empty((wl.batch, ev.ncons), dtype=torch.float64, device=dev)
            jac = torch.empty((wl.batch, ev.nnz), dtype=torch.float64, device=dev)
            ev.eval_ptr(x.data_ptr(), f.data_ptr(), g.data_ptr(), jac.data_ptr(), mode, capi.MEM_DEVICE, None)
            ev.sync()
            assert np.array_equal(got["f"], f.cpu().numpy()) and np.array_equal(got["g"], g.cpu().numpy())
            assert np.array_equal(got["jac"], jac.cpu().numpy()), (name, mode, chunks)
            assert np.isfinite(got["grad"]).all()
        only_g = ev.eval_host(wl.x, want=("g",), jac_mode=W.JAC_FD)
        assert only_g["f"] is None and only_g["jac"] is None and np.array_equal(only_g["g"], got["g"])
        ev.close()


def test_full_size_batch_properties():
    """BASELINE config C2 at full size (4096 instances): properties that need no 4096-instance oracle run."""
    import torch
    wl = W.pm3d(batch=4096)
    ev = capi.Evaluator(wl)
    dev = torch.device("cuda:0")
    x = torch.from_numpy(wl.x).to(dev)
    out = {}
    for mode in (W.JAC_FD, W.JAC_EXACT):
        f = torch.empty(wl.batch, dtype=torch.float64, device=dev)
        g = torch.empty((wl.batch, ev.ncons), dtype=torch.float64, device=dev)
        jac = torch.full((wl.batch, ev.nnz), float("nan"), dtype=torch.float64, device=dev)
        ev.eval_ptr(x.data_ptr(), f.data_ptr(), g.data_ptr(), jac.data_ptr(), mode, capi.MEM_DEVICE, None)
        ev.sync()
        out[mode] = (f.cpu().numpy(), g.cpu().numpy(), jac.cpu().numpy())
    f, g, jfd = out[W.JAC_FD]
    _, _, jex = out[W.JAC_EXACT]
    assert not np.isnan(jfd).any() and not np.isnan(jex).any()
    # (1) FD ~ exact on every instance. The bound is the FD noise model eps*|g|/(2*delta): with
    # PSOPT's step 2^-26*(1+|z|) a position clipped to ~0 and |g| ~ 1e6 m^2 gives ~7e-3 absolute.
    assert np.abs(jfd - jex).max() <= 5e-5 * np.abs(jex).max()
    # (2) a random sample of instances is bit-identical to the oracle
    idx = np.random.default_rng(3).choice(wl.batch, 12, replace=False)
    orc = ob.Oracle(wl)
    for b in idx:
        ref = orc.eval(wl.x[b:b + 1], jac_mode=W.JAC_FD, first=int(b), count=1)
        assert np.array_equal(ref["f"][0], f[b]) and np.array_equal(ref["g"][0], g[b])
        assert np.array_equal(ref["jac"][0], jfd[b])
    # (3) instances are independent: a permuted batch gives the permuted result
    perm = np.random.default_rng(4).permutation(wl.batch)
    wl2 = W.pm3d(batch=4096)
    wl2.cylinders, wl2.x = wl.cylinders[perm], wl.x[perm]
    ev2 = capi.Evaluator(wl2)
    got = ev2.eval_host(wl2.x, jac_mode=W.JAC_FD)
    assert np.array_equal(got["g"], g[perm]) and np.array_equal(got["jac"], jfd[perm])
    # (4) the duration row equals tf - t0 and the event rows equal the boundary states
    N = 40
    assert np.array_equal(g[:, -1], wl.x[:, wl.itf(0)] - wl.x[:, wl.it0(0)])
    assert np.array_equal(g[:, 6 * N:6 * N + 6], wl.x[:, wl.ix(0, 0, 0):wl.ix(0, 0, 0) + 6])
    ev.close()
    ev2.close()


def test_summary_matches_numpy(evaluators):
    ev, orc, wl = _get(evaluators, "C2-pm3d-64")
    s = ev.summary_host(wl.x)
    ref = ev.eval_host(wl.x, want=("f", "g"))
    viol = np.maximum(np.maximum(wl.gl - ref["g"], ref["g"] - wl.gu), 0.0).max(axis=1)
    assert np.array_equal(s[:, 0], ref["f"]) and np.array_equal(s[:, 1], viol)


def test_ipopt_shaped_shims():
    wl = W.reference_vgp("ocp")
    ev = capi.Evaluator(wl)
    orc = ob.Oracle(wl)
    L, h = ev.L, ev.h
    n, m, nz = ev.nvars, ev.ncons, ev.nnz
    x = np.ascontiguousarray(wl.x[0])
    dp = C.POINTER(C.c_double)
    ip = C.POINTER(C.c_int32)
    obj = C.c_double()
    assert L.ecuda_ipopt_eval_f(h, n, x.ctypes.data_as(dp), 1, C.byref(obj)) == 0
    g = np.zeros(m)
    assert L.ecuda_ipopt_eval_g(h, n, x.ctypes.data_as(dp), 0, m, g.ctypes.data_as(dp)) == 0
    grad = np.zeros(n)
    assert L.ecuda_ipopt_eval_grad_f(h, n, x.ctypes.data_as(dp), 0, grad.ctypes.data_as(dp)) == 0
    irow, jcol = np.zeros(nz, dtype=np.int32), np.zeros(nz, dtype=np.int32)
    assert L.ecuda_ipopt_eval_jac_g(h, n, x.ctypes.data_as(dp), 0, m, nz, irow.ctypes.data_as(ip),
                                    jcol.ctypes.data_as(ip), None) == 0
    vals = np.zeros(nz)
    assert L.ecuda_ipopt_eval_jac_g(h, n, x.ctypes.data_as(dp), 0, m, nz, None, None, vals.ctypes.data_as(dp)) == 0
    ref = orc.eval(wl.x, want=("f", "g", "jac", "grad"), jac_mode=W.JAC_EXACT)
    oi, oj, _ = orc.structure()
    assert obj.value == ref["f"][0] and np.array_equal(g, ref["g"][0])
    assert rel_err(grad, ref["grad"][0]) <= TOL_VALUE and rel_err(vals, ref["jac"][0]) <= TOL_JAC
    assert np.array_equal(irow, oi) and np.array_equal(jcol, oj)
    assert L.ecuda_ipopt_eval_g(h, n + 1, x.ctypes.data_as(dp), 0, m, g.ctypes.data_as(dp)) == -1
    ev.close()


def test_error_behaviour_on_gpu():
    wl = W.pm3d(batch=2)
    ev = capi.Evaluator(wl, upload=False)
    with pytest.raises(capi.EcudaError, match="upload_instances"):
        ev.eval_host(wl.x)
    h = C.c_void_p()
    assert capi.lib().ecuda_create(999, C.byref(h)) == -1
    ev.close()


def test_injected_collocation_is_used():
    """ecuda_set_collocation: a perturbed D must change the defects exactly as the oracle formula says"""
    wl = W.pm3d(batch=1, nnodes=9, ncyl=0)
    ev = capi.Evaluator(wl)
    tau, w, D = ev.collocation(0)
    base = ev.eval_host(wl.x, want=("g",))["g"]
    ev.set_collocation(0, tau, w, 2.0 * D)
    twice = ev.eval_host(wl.x, want=("g",))["g"]
    N, ns = 9, 6
    X = np.stack([[wl.x[0, wl.ix(0, k, i)] for i in range(ns)] for k in range(N)])
    assert np.allclose((twice - base)[0, :N * ns].reshape(N, ns), D @ X, rtol=1e-12, atol=1e-9)
    ev.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["C0-ocp", "C0-mip", "C2-pm3d-64", "C4-multiphase", "C2-pm3d-scaled-deps"])
def test_fast_and_generic_kernels_agree_bitwise(name):
    """the specialised kernels (k_eval_fast) and the generic ones (k_eval) must produce the same bits:
    both are built from the same operation sequences, only the work layout differs"""
    wl = CASES[name]()
    out = {}
    # "fast": k_eval_fast, exact mode streams the template with the TMA copy warp (the default);
    # "fast-image": exact through the persistent k_eval_image (shared-memory image + bulk stores);
    # "fast-ldst": template copied with plain loads/stores; "generic": k_eval
    # ("fast" is k_eval_rows, the row-owner kernel; "columns" the column-owner k_eval_fast)
    # round 2: "fast" now means the N-specialised k_rows_n for finite differences wherever an instantiation exists
    # (all five cases); "rows-r1" switches it off (round-1 k_eval_rows); "ring" / "stream" select the two other
    # exact-mode kernels (k_rows_n with the shared-memory store ring, k_stream_exact); "persist" the persistent
    # finite-difference kernel with TMA prefetch of the next instance (k_rows_n_fd_persist)
    for tag, env in (("fast", {}), ("fast-image", {"ECUDA_IMAGE": "1"}), ("fast-ldst", {"ECUDA_NO_COPY_WARP": "1"}),
                     ("columns", {"ECUDA_NO_ROWS": "1"}), ("generic", {"ECUDA_NO_FAST": "1"}),
                     ("rows-r1", {"ECUDA_NO_ROWSN": "1"}), ("ring", {"ECUDA_EXACT_KERNEL": "ring"}),
                     ("stream", {"ECUDA_EXACT_KERNEL": "stream"}), ("persist", {"ECUDA_PERSIST": "1"})):
        os.environ.update(env)  # read by ecuda_create
        try:
            ev = capi.Evaluator(wl, device=0)
        finally:
            for k in env:
                os.environ.pop(k, None)
        out[tag] = {m: ev.eval_host(wl.x, want=("f", "g", "jac"), jac_mode=m) for m in (W.JAC_FD, W.JAC_EXACT)}
        ev.close()
    for m in (W.JAC_FD, W.JAC_EXACT):
        for key in ("f", "g", "jac"):
            assert np.array_equal(out["fast"][m][key], out["generic"][m][key]), (name, m, key)
            assert np.array_equal(out["fast-ldst"][m][key], out["generic"][m][key]), (name, m, key)
            assert np.array_equal(out["fast-image"][m][key], out["generic"][m][key]), (name, m, key)
            assert np.array_equal(out["columns"][m][key], out["generic"][m][key]), (name, m, key)
            for tag in ("rows-r1", "ring", "stream", "persist"):
                assert np.array_equal(out[tag][m][key], out["generic"][m][key]), (name, m, key, tag)


def test_peer_barrier_reports_an_absent_peer(monkeypatch):
    """a peer that never arrives must be detectable through the ABI. One GPU plays rank 0 of 2; rank 1's
    flag slot is never written, so the bounded wait expires: ecuda_peer_barrier_status says so (step, late rank),
    ecuda_sync returns ECUDA_ERR_PEER, and the status is sticky until it is read with reset."""
    import torch
    monkeypatch.setenv("ECUDA_PEER_TIMEOUT_MS", "20")
    wl = W.pm3d(batch=2, nnodes=9, ncyl=1)
    ev = capi.Evaluator(wl, device=0)
    dev = torch.device("cuda:0")
    stream = torch.cuda.Stream(device=dev)
    st = stream.cuda_stream
    flags = torch.zeros(64, dtype=torch.int64, device=dev)
    ghost = torch.zeros(64, dtype=torch.int64, device=dev)  # stands for the absent rank's array
    torch.cuda.synchronize()
    # a barrier every rank reaches: one rank, completes at once
    ev.peer_barrier_ptr([flags.data_ptr()], 0, 1, st)
    assert ev.peer_barrier_status(st) == (False, 0, -1)
    assert ev.sync_status()[0] == 0
    # two ranks, the second never shows up
    ev.peer_barrier_ptr([flags.data_ptr(), ghost.data_ptr()], 0, 2, st)
    timed_out, step, late = ev.peer_barrier_status(st)
    assert (timed_out, step, late) == (True, 2, 1)
    rc, msg = ev.sync_status()
    assert rc == -5 and "timed out at step 2" in msg
    assert ev.peer_barrier_status(st)[0] is True                       # sticky
    assert ev.peer_barrier_status(st, reset=True) == (True, 2, 1)      # read and clear
    assert ev.peer_barrier_status(st) == (False, 0, -1)
    assert ev.sync_status()[0] == 0
    assert int(ghost[0].item()) == 2  # this rank did publish its arrival to the peer's array
    ev.close()
