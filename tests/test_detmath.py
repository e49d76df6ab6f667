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
