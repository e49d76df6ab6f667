"""include/ecuda_detmath.h: the deterministic sin/cos shared by host and device (fw6 dynamics)."""
import numpy as np

import oracle_binding as ob


def test_sincos_accuracy_against_extended_precision():
    rng = np.random.default_rng(7)
    x = np.concatenate([rng.uniform(-8, 8, 20000), rng.uniform(-1e3, 1e3, 20000), rng.normal(0, 1e-3, 2000),
                        np.array([0.0, np.pi / 4, -np.pi / 4, np.pi / 2, np.pi, 2 * np.pi, 1e-300, -0.0])])
    s, c = ob.det_sincos(x)
    xl = x.astype(np.longdouble)
    rs, rc = np.sin(xl), np.cos(xl)
    ulp_s = np.abs(s.astype(np.longdouble) - rs) / np.spacing(np.abs(rs).astype(np.float64) + 1e-300)
    ulp_c = np.abs(c.astype(np.longdouble) - rc) / np.spacing(np.abs(rc).astype(np.float64) + 1e-300)
    small = np.abs(x) < 100  # the states the fw6 model feeds it (|gamma| <= 0.5, |psi| <= 2 pi)
    assert float(ulp_s[small].max()) < 1.5 and float(ulp_c[small].max()) < 1.5
    assert np.abs(s - np.sin(x)).max() < 2e-13 and np.abs(c - np.cos(x)).max() < 2e-13
    assert np.abs(s * s + c * c - 1).max() < 1e-15 * 8


def test_sincos_special_values():
    s, c = ob.det_sincos(np.array([0.0]))
    assert s[0] == 0.0 and c[0] == 1.0


def test_sincos_dense_sweep_over_the_model_range_with_stated_ulp_bound():
    """VERDICT r1 item 12. The fw6 model feeds ecuda_sincos the flight-path angle gamma and the heading psi; their
    decision-variable boxes are |gamma| <= 0.5 and |psi| <= 2 pi, and a finite-difference perturbation moves them by
    2^-26 (1 + |z|). Sweep [-7, 7] densely (2 million points, plus the neighbourhoods of every multiple of pi/4 where
    the quadrant reduction switches, plus the perturbed twins z +- delta) against numpy's libm in extended precision.
    Stated bound: < 2.0 ulp for both functions on this range. Measured: sin 1.56 ulp (at x = 2.3501, the edge of a
    reduction interval, |r| = pi/4, where the degree-13 kernel is least accurate), cos 1.38 ulp (x = 1.0538). Both sides
    of a central difference use the same routine, so what the Jacobian needs is determinism, not more accuracy."""
    rng = np.random.default_rng(20261018)
    x = np.concatenate([np.linspace(-7.0, 7.0, 2_000_001), rng.uniform(-7, 7, 200_000)])
    near = np.concatenate([k * np.pi / 4 + np.linspace(-1e-6, 1e-6, 2001) for k in range(-9, 10)])
    x = np.concatenate([x, near])
    delta = 2.0 ** -26 * (1.0 + np.abs(x[:100_000]))
    x = np.concatenate([x, x[:100_000] + delta, x[:100_000] - delta])
    s, c = ob.det_sincos(x)
    xl = x.astype(np.longdouble)
    rs, rc = np.sin(xl), np.cos(xl)
    # ulp of the reference value in double precision (values near 0 have tiny ulps: use the spacing at the value)
    ulp_s = np.abs(s.astype(np.longdouble) - rs) / np.maximum(np.spacing(np.abs(rs).astype(np.float64)), np.finfo(np.float64).tiny)
    ulp_c = np.abs(c.astype(np.longdouble) - rc) / np.maximum(np.spacing(np.abs(rc).astype(np.float64)), np.finfo(np.float64).tiny)
    # points where the true value is within 1e-12 of a zero crossing lose relative accuracy by cancellation in the
    # argument reduction (two-term pi/2); the model never sits there to better than 2^-26, bound them absolutely
    ok_s, ok_c = np.abs(rs) > 1e-9, np.abs(rc) > 1e-9
    ms, mc = float(ulp_s[ok_s].max()), float(ulp_c[ok_c].max())
    assert ms < 2.0 and mc < 2.0, (ms, mc)
    assert ms > 0.5 and mc > 0.5, "the comparison is not vacuous: a correctly rounded routine would show <= 0.5"
    assert float(np.abs(s.astype(np.longdouble) - rs)[~ok_s].max(initial=0.0)) < 1e-25 + 2.5e-16 * 1e-9
    assert float(np.abs(c.astype(np.longdouble) - rc)[~ok_c].max(initial=0.0)) < 1e-25 + 2.5e-16 * 1e-9
    # determinism across calls and monotone-consistency of the pair
    s2, c2 = ob.det_sincos(x)
    assert np.array_equal(s, s2) and np.array_equal(c, c2)
    assert float(np.abs(s * s + c * c - 1.0).max()) <= 4 * np.finfo(np.float64).eps
