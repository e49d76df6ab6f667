"""Mesh-refinement support: the interpolation data behind ecuda_ode_error /
ecuda_resample (host helpers of libecuda.so, no GPU) and the oracle's restatement of the error estimate.
The device kernels are compared with the oracle in tests/test_gpu_parity.py."""
import numpy as np
import pytest

import oracle_binding as ob
from etol_b200 import capi, workloads as W


def _mesh(kind, N):
    Q = 4
    tq, wq = np.zeros((N - 1) * Q), np.zeros((N - 1) * Q)
    E, dE = np.zeros(((N - 1) * Q, N)), np.zeros(((N - 1) * Q, N))
    p = lambda a: a.ctypes.data_as(capi._dp)
    assert capi.lib().ecuda_host_error_mesh(kind, N, p(tq), p(wq), p(E), p(dE)) == 0
    return tq, wq, E, dE


@pytest.mark.parametrize("kind", [W.LEGENDRE, W.CHEBYSHEV])
@pytest.mark.parametrize("N", [2, 5, 17, 40, 120])
def test_error_mesh_is_an_exact_interpolation_rule(kind, N):
    tau, _, D = capi.host_collocation(kind, N)
    tq, wq, E, dE = _mesh(kind, N)
    assert np.all(tq > -1) and np.all(tq < 1) and np.all(np.diff(tq) > 0)
    assert abs(wq.sum() - 2.0) < 1e-13                       # the intervals tile [-1, 1]
    assert np.abs(E.sum(axis=1) - 1.0).max() < 1e-10         # partition of unity
    assert np.abs(dE.sum(axis=1)).max() < 1e-7 * N * N
    deg = min(N - 1, 6)                                      # polynomials of degree < N are reproduced
    c = np.arange(1, deg + 2, dtype=np.float64)
    poly = np.polynomial.Polynomial(c)
    assert np.abs(E @ poly(tau) - poly(tq)).max() < 1e-9 * np.abs(poly(tq)).max()
    assert np.abs(dE @ poly(tau) - poly.deriv()(tq)).max() < 1e-8 * N * N * max(1.0, np.abs(poly.deriv()(tq)).max())


@pytest.mark.parametrize("kind", [W.LEGENDRE, W.CHEBYSHEV])
def test_resample_matrix(kind):
    R = np.zeros((12, 12))
    assert capi.lib().ecuda_host_resample_matrix(kind, 12, 12, R.ctypes.data_as(capi._dp)) == 0
    assert np.array_equal(R, np.eye(12))                     # same mesh: identity, exactly
    R = np.zeros((31, 12))
    assert capi.lib().ecuda_host_resample_matrix(kind, 12, 31, R.ctypes.data_as(capi._dp)) == 0
    t12, t31 = capi.host_collocation(kind, 12)[0], capi.host_collocation(kind, 31)[0]
    poly = np.polynomial.Polynomial([0.3, -1.0, 2.0, 0.5, -0.25])
    assert np.abs(R @ poly(t12) - poly(t31)).max() < 1e-12
    assert np.array_equal(R[0], np.eye(12)[0]) and np.array_equal(R[-1], np.eye(12)[-1])  # end points are nodes


def test_oracle_error_known_answer_single_integrator():
    """x_i(tau) = tau^2, u = 0 on [t0, tf] = [-1, 1]: x' - u = 2 tau, so eta_k = |tau_k+1^2 - tau_k^2| on every
    interval that does not contain 0, w_i = max(|x|, |x'|) = 2, eps_k = eta_k / 3"""
    wl = W.reference_vgp("ocp")
    o = ob.Oracle(wl)
    N = wl.nnodes[0]
    tau = capi.host_collocation(wl.collocation, N)[0]
    z = np.zeros((1, wl.nvars))
    for k in range(N):
        z[0, wl.ix(0, k, 0)] = z[0, wl.ix(0, k, 1)] = tau[k] ** 2
    z[0, wl.it0(0)], z[0, wl.itf(0)] = -1.0, 1.0
    err = ob.ode_error(o, wl, z)[0]
    want = np.abs(tau[1:] ** 2 - tau[:-1] ** 2) / 3.0
    away = (tau[1:] * tau[:-1]) > 0
    assert away.sum() >= N - 3
    assert np.abs(err[away] - want[away]).max() < 1e-13


def test_oracle_error_vanishes_on_an_exact_trajectory_and_not_elsewhere():
    """pm3d under constant acceleration: position quadratic, velocity linear in t -- representable exactly"""
    wl = W.pm3d(batch=1, nnodes=12, ncyl=0)
    o = ob.Oracle(wl)
    N = wl.nnodes[0]
    tau = capi.host_collocation(wl.collocation, N)[0]
    t0, tf = 0.0, 8.0
    t = 0.5 * (tf - t0) * tau + 0.5 * (tf + t0)
    a, v0, p0 = np.array([0.5, -0.25, 0.125]), np.array([1.0, 2.0, -1.0]), np.array([3.0, 4.0, 5.0])
    z = np.zeros((1, wl.nvars))
    for k in range(N):
        for i in range(3):
            z[0, wl.ix(0, k, i)] = p0[i] + v0[i] * t[k] + 0.5 * a[i] * t[k] ** 2
            z[0, wl.ix(0, k, 3 + i)] = v0[i] + a[i] * t[k]
            z[0, wl.iu(0, k, i)] = a[i]
    z[0, wl.it0(0)], z[0, wl.itf(0)] = t0, tf
    assert ob.ode_error(o, wl, z).max() < 1e-13
    z[0, wl.iu(0, N // 2, 0)] += 1.0                          # a kink in the control: the defect shows up
    err = ob.ode_error(o, wl, z)[0]
    assert err.max() > 1e-3 and err.argmax() in (N // 2 - 1, N // 2)


def test_oracle_resample_round_trip_and_scaling():
    wl = W.pm3d_multiphase(batch=2, nphases=2, nnodes=9, ncyl=1, scaled=True)
    o = ob.Oracle(wl)
    up = ob.resample(o, wl, wl.x, [21, 17])                   # unscaled values on the finer meshes
    fine = W.pm3d_multiphase(batch=2, nphases=2, nnodes=9, ncyl=1)
    fine.nnodes = [21, 17]
    assert up.shape == (2, fine.nvars)
    z = wl.x / wl.sz
    for p in range(2):                                        # end nodes and times carry over
        for i in range(6):
            assert np.allclose(up[:, fine.ix(p, 0, i)], z[:, wl.ix(p, 0, i)], rtol=0, atol=1e-9)
            assert np.allclose(up[:, fine.ix(p, fine.nnodes[p] - 1, i)], z[:, wl.ix(p, 8, i)], rtol=0, atol=1e-9)
        assert np.array_equal(up[:, fine.itf(p)], z[:, wl.itf(p)])
    o2 = ob.Oracle(fine)                                      # and back: a degree-8 polynomial survives the detour
    back = ob.resample(o2, fine, up, [9, 9], sz_new=wl.sz)
    assert np.abs(back - wl.x).max() < 1e-9 * np.abs(wl.x).max()
