import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def rel_err(a, b):
    """max_i |a_i-b_i| / max(|a_i|,|b_i|), exact zeros on both sides count as 0."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    d = np.abs(a - b)
    den = np.maximum(np.abs(a), np.abs(b))
    with np.errstate(invalid="ignore", divide="ignore"):
        r = np.where(d == 0.0, 0.0, d / den)
    if np.isnan(a).any() or np.isnan(b).any():
        return float("inf")
    return float(r.max()) if r.size else 0.0


# tolerances of BASELINE.json north_star
TOL_VALUE = 1e-12   # objective / constraint values, relative
TOL_JAC = 1e-9      # Jacobian entries, relative, same perturbation step


@pytest.fixture(scope="session")
def oracle_lib():
    import oracle_binding
    return oracle_binding.lib()
