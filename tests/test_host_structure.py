"""Host half of libecuda.so (no GPU needed): dimensions, collocation data, sparsity pattern, CPR
column groups and obstacle-edge records must equal the oracle's bit for bit (north_star: "bit-exact
sparsity pattern and colouring")."""
import ctypes as C

import numpy as np
import pytest

import oracle_binding as ob
from etol_b200 import capi, workloads as W

CASES = [
    lambda: W.reference_vgp("ocp"), lambda: W.reference_vgp("mip"),
    lambda: W.reference_vgp("ocp", pattern_mode=W.MODEL_DEPS, index_base=1),
    lambda: W.pm3d(batch=1), lambda: W.pm3d(batch=1, pattern_mode=W.MODEL_DEPS),
    lambda: W.fw6(batch=1), lambda: W.fw6(batch=1, nnodes=50, ncyl=7, pattern_mode=W.MODEL_DEPS),
    lambda: W.pm3d_multiphase(batch=1), lambda: W.pm3d_multiphase(batch=1, nphases=2, nnodes=6, ncyl=0),
    lambda: W.pm3d(batch=1, nnodes=2, ncyl=1),
]


@pytest.mark.parametrize("mk", CASES)
def test_structure_matches_oracle_bitwise(mk):
    wl = mk()
    o = ob.Oracle(wl)
    d = capi.host_dims(wl)
    assert (d.nvars, d.ncons, d.nnz, d.ngroups) == (o.nvars, o.ncons, o.nnz, o.ngroups)
    assert (d.nstates, d.ncontrols, d.nlinkages) == (o.ns, o.nc, o.nlink)
    irow, jcol, grp = capi.host_structure(wl)
    oi, oj, og = o.structure()
    assert np.array_equal(irow, oi) and np.array_equal(jcol, oj) and np.array_equal(grp, og)
    base = wl.index_base
    # sorted by (col,row), no duplicates, indices in range
    key = (jcol.astype(np.int64) - base) * d.ncons + (irow - base)
    assert np.all(np.diff(key) > 0)
    assert irow.min() >= base and irow.max() < d.ncons + base and jcol.max() < d.nvars + base
    # a valid CPR grouping: columns of one group never share a row
    for g in range(d.ngroups):
        rows = np.concatenate([irow[jcol == c + base] for c in np.nonzero(grp == g)[0]])
        assert rows.size == np.unique(rows).size


@pytest.mark.parametrize("kind", [W.LEGENDRE, W.CHEBYSHEV])
@pytest.mark.parametrize("N", [2, 3, 5, 17, 30, 33, 40, 41, 200])
def test_collocation_matches_oracle_bitwise(kind, N):
    a = capi.host_collocation(kind, N)
    b = ob.make_collocation(kind, N)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_edge_records_match_reference_geometry():
    for poly in W.REF_BORDERS:
        rec = capi.edge_records(poly[:, :2])
        n = len(poly)
        for e in range(n):
            xc, yc, radsq, tt, asq, bsq = ob.edge_geometry(poly[e], poly[(e + 1) % n])
            assert list(rec[e]) == [xc, yc, np.cos(tt), np.sin(tt), asq, bsq]


def test_bad_descriptions_are_rejected():
    wl = W.pm3d(batch=1)
    for field, val in (("model", 7), ("nphases", 0), ("nphases", 9), ("batch", 0), ("index_base", 2),
                       ("pattern_mode", 5), ("collocation", 3), ("ncontrols", 5), ("ntracks", 1)):
        d = capi.make_desc(wl)
        setattr(d, field, val)
        assert capi.lib().ecuda_host_dims(C.byref(d), C.byref(capi.Dims())) == -1, field
    d = capi.make_desc(wl)
    d.nnodes[0] = 1
    assert capi.lib().ecuda_host_dims(C.byref(d), C.byref(capi.Dims())) == -1


def test_pack_instances_layout():
    wl = W.reference_vgp("ocp")
    d = capi.host_dims(wl)
    inst = capi.pack_instances(wl, d)
    assert d.rec_size == 6 and d.track_size == 7 and d.inst_stride % 4 == 0
    assert inst.shape == (1, d.inst_stride) and d.inst_stride >= 9 * 6 + 2 * 7
    assert inst[0, 54] == 0.5 and list(inst[0, 55:61]) == [0.0, 1.51, 2.0, 32.0, 2.0, 2.0]
    wl = W.pm3d_multiphase(batch=3, nnodes=5, ncyl=2)
    d = capi.host_dims(wl)
    inst = capi.pack_instances(wl, d)
    assert d.rec_size == 4 and inst.shape == (3, 24)
    assert np.array_equal(inst[:, 2::4], wl.cylinders[:, :, 2] ** 2)


@pytest.mark.parametrize("mk", [lambda: W.reference_vgp("ocp", batch=3, jitter=0.02), lambda: W.pm3d(batch=4),
                                lambda: W.pm3d(batch=3, scaled=True, pattern_mode=W.MODEL_DEPS),
                                lambda: W.pm3d_multiphase(batch=3, scaled=True), lambda: W.fw6(batch=2, nnodes=21, ncyl=5),
                                lambda: W.pm3d(batch=2, nnodes=2, ncyl=1)])
def test_compact_split_of_the_exact_jacobian(mk):
    """ecuda_eval_compact's split: the triplets outside local_index are the D-coupled
    off-diagonal entries; in the oracle's exact Jacobian they are the same for every instance (different x, different
    obstacles), and splicing the per-instance part back with ecuda_splice_jacobian gives the full array bit for bit."""
    wl = mk()
    o = ob.Oracle(wl)
    idx = capi.host_compact_structure(wl)
    d = capi.host_dims(wl)
    irow, jcol, _ = capi.host_structure(wl)
    nshared = sum(N * (N - 1) * d.nstates for N in wl.nnodes)
    assert idx.size == d.nnz - nshared and np.all(np.diff(idx) > 0)
    jac = o.eval(wl.x, want=("jac",), jac_mode=W.JAC_EXACT)["jac"]
    shared_mask = np.ones(d.nnz, dtype=bool)
    shared_mask[idx] = False
    assert np.all(jac[:, shared_mask] == jac[0, shared_mask])      # instance-independent
    assert np.all(jac[0, shared_mask] != 0.0)                      # and structurally non-zero (D has no zeros off the diagonal)
    shared = np.where(shared_mask, jac[0], 0.0)
    full = capi.splice_jacobian(shared, idx, jac[:, idx], d.nnz)
    assert np.array_equal(full, jac)
    # bad arguments are refused
    bad = idx.copy()
    bad[0] = d.nnz
    with pytest.raises(capi.EcudaError):
        capi.splice_jacobian(shared, bad, jac[:, idx], d.nnz)
