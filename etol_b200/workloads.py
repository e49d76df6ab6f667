"""Synthetic VGP workloads for the eCUDA evaluator.

A Workload is the *raw* vehicle-guidance-problem data of a batch of independent instances, in the
shape the reference example holds it (polygon corners, track waypoint tables, cylinders), plus the
NLP description, bounds and a batch of fixed decision vectors. The same Workload feeds the CUDA
path (after `pack_instances`) and the CPU oracle (tests only), so both see identical inputs.

C0 restates the shipped reference VGP (resource/configs/ocp_2d_ex1.xml:2-47 of the ETOL tree) and
its `mip_2d_ex1.xml` variant; C1-C5 are the build-defined extensions of SURVEY.md section 8(d).
"""
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

SI2D, PM3D, FW6 = 0, 1, 2
LEGENDRE, CHEBYSHEV = 0, 1
DENSE_NODE, MODEL_DEPS = 0, 1
JAC_EXACT, JAC_FD = 0, 1

MODEL_SHAPE = {SI2D: (2, 2), PM3D: (6, 3), FW6: (6, 3)}  # (nstates, default ncontrols)
SEED = 0xE701


@dataclass
class Workload:
    name: str
    model: int
    nnodes: List[int]
    nstatic: List[int]
    batch: int
    ncontrols: int = 0
    ntracks: int = 0
    nwaypoints: int = 0
    collocation: int = LEGENDRE
    pattern_mode: int = DENSE_NODE
    maximize: bool = False
    index_base: int = 0
    # raw VGP data
    borders: Optional[list] = None    # [B][nphases] -> list of (ncorners,3) arrays   (si2d)
    tracks: Optional[list] = None     # [B] -> list of (radius, t[], x[], y[])           (si2d)
    cylinders: Optional[np.ndarray] = None  # [B][sum nstatic][3] = cx, cy, r            (pm3d/fw6)
    # scaling (shared by the batch); None = ones
    sz: Optional[np.ndarray] = None
    sg: Optional[np.ndarray] = None
    sf: float = 1.0
    # unscaled variable bounds (shared) and constraint bounds per instance
    zl: Optional[np.ndarray] = None
    zu: Optional[np.ndarray] = None
    gl: Optional[np.ndarray] = None
    gu: Optional[np.ndarray] = None
    x: Optional[np.ndarray] = None    # [B][nvars] scaled decision vectors
    meta: dict = field(default_factory=dict)
    tape: Optional[object] = None     # etol_b200.tape.Tape of a user model (model = its registered id)

    @property
    def nphases(self):
        return len(self.nnodes)

    @property
    def ns(self):
        return self.tape.ns if self.tape is not None else MODEL_SHAPE[self.model][0]

    @property
    def nc(self):
        if self.tape is not None:
            return self.tape.nc
        return self.ncontrols if self.ncontrols > 0 else MODEL_SHAPE[self.model][1]

    @property
    def npath(self):
        nuser = len(getattr(self.tape, "row_out", [])) if getattr(self, "tape", None) is not None else 0
        return [s + self.ntracks + nuser for s in self.nstatic]

    @property
    def nvars(self):
        return sum((self.ns + self.nc) * n + 2 for n in self.nnodes)

    @property
    def ncons(self):
        per = sum(self.ns * n + 2 * self.ns + q * n + 1 for n, q in zip(self.nnodes, self.npath))
        return per + (self.nphases - 1) * (self.ns + 1)

    # ---- layout helpers (mirror include/ecuda.h) -------------------------------------------------
    def zoff(self, p):
        return sum((self.ns + self.nc) * n + 2 for n in self.nnodes[:p])

    def goff(self, p):
        return sum(self.ns * n + 2 * self.ns + q * n + 1 for n, q in zip(self.nnodes[:p], self.npath[:p]))

    def iu(self, p, k, j):
        return self.zoff(p) + k * self.nc + j

    def ix(self, p, k, i):
        return self.zoff(p) + self.nc * self.nnodes[p] + k * self.ns + i

    def it0(self, p):
        return self.zoff(p) + (self.ns + self.nc) * self.nnodes[p]

    def itf(self, p):
        return self.it0(p) + 1

    def slice_batch(self, lo, hi):
        """Contiguous instance range [lo,hi) as its own Workload (multi-GPU sharding)."""
        import copy
        w = copy.copy(self)
        w.batch = hi - lo
        w.borders = None if self.borders is None else self.borders[lo:hi]
        w.tracks = None if self.tracks is None else self.tracks[lo:hi]
        w.cylinders = None if self.cylinders is None else np.ascontiguousarray(self.cylinders[lo:hi])
        w.gl = None if self.gl is None else np.ascontiguousarray(self.gl[lo:hi])
        w.gu = None if self.gu is None else np.ascontiguousarray(self.gu[lo:hi])
        w.x = None if self.x is None else np.ascontiguousarray(self.x[lo:hi])
        return w


def _rng(seed):
    return np.random.Generator(np.random.Philox(key=int(seed)))


def _decision_vectors(wl, guess, rng, spread=0.05):
    """x = guess + spread*U(-1,1)*(ub-lb), clipped strictly inside the bounds, then scaled."""
    B, n = wl.batch, wl.nvars
    span = wl.zu - wl.zl
    x = guess + spread * rng.uniform(-1.0, 1.0, size=(B, n)) * span
    eps = 1e-6 * span
    x = np.minimum(np.maximum(x, wl.zl + eps), wl.zu - eps)
    fixed = span == 0.0
    x[:, fixed] = wl.zl[fixed]
    if wl.sz is not None:
        x = x * wl.sz
    return np.ascontiguousarray(x)


def psopt_like_scaling(wl):
    """PSOPT scaling="automatic" rule for variables: 1/max(|lb|,|ub|)."""
    m = np.maximum(np.abs(wl.zl), np.abs(wl.zu))
    sz = np.where((m > 0) & np.isfinite(m), 1.0 / np.where(m > 0, m, 1.0), 1.0)
    return sz


# ---- C0: the reference VGP -------------------------------------------------------------------------
# Values of resource/configs/ocp_2d_ex1.xml:2-47 (and the mip_2d_ex1.xml differences).
REF_BORDERS = [
    np.array([[3.20, 2.50, 0.0], [3.40, 2.60, 0.0], [3.50, 3.40, 0.0], [3.30, 3.00, 0.0], [3.10, 3.50, 0.0]]),
    np.array([[2.20, 2.50, 0.0], [2.40, 2.60, 0.0], [2.50, 3.40, 0.0], [2.10, 3.50, 0.0]]),
]
REF_TRACKS = {
    "ocp": [(0.50, [0.0, 32.0], [1.51, 2.00], [2.00, 2.00]), (0.50, [0.0, 32.0], [1.00, 1.00], [4.00, 3.00])],
    "mip": [(0.50, [0.0, 32.0], [2.00, 2.50], [2.00, 2.00]), (0.50, [0.0, 32.0], [1.00, 1.00], [4.00, 3.00])],
}


def reference_vgp(variant="ocp", batch=1, jitter=0.0, seed=SEED, **kw):
    """C0. variant "ocp": nsteps=32, 2 controls (134/434/3372); "mip": nsteps=16, 4 controls."""
    nsteps, nc = (32, 2) if variant == "ocp" else (16, 4)
    dt = 0.5
    N = nsteps + 1
    rng = _rng(seed)
    nedges = sum(len(b) for b in REF_BORDERS)
    wl = Workload(name=f"C0-si2d-{variant}", model=SI2D, nnodes=[N], nstatic=[nedges], batch=batch,
                  ncontrols=nc, ntracks=2, nwaypoints=2, **kw)
    borders, tracks = [], []
    for b in range(batch):
        polys = []
        for poly in REF_BORDERS:
            q = poly.copy()
            if jitter and b > 0:
                q[:, :2] += jitter * rng.uniform(-1, 1, size=(len(poly), 2))
            polys.append(q)
        borders.append([polys])
        trs = []
        for (r, t, x, y) in REF_TRACKS[variant]:
            x, y = np.array(x, dtype=np.float64), np.array(y, dtype=np.float64)
            if jitter and b > 0:
                x = x + jitter * rng.uniform(-1, 1, size=2)
                y = y + jitter * rng.uniform(-1, 1, size=2)
            trs.append((r, np.array(t, dtype=np.float64), x, y))
        tracks.append(trs)
    wl.borders, wl.tracks = borders, tracks
    # bounds, ePSOPT::addBounds (src/ePSOPT/ePSOPT.cpp:125-155)
    x0, xf, xtol = np.array([1.0, 2.0]), np.array([5.0, 4.0]), np.array([0.01, 0.01])
    tspan = nsteps * dt
    zl, zu = np.zeros(wl.nvars), np.zeros(wl.nvars)
    for k in range(N):
        for j in range(nc):
            zl[wl.iu(0, k, j)], zu[wl.iu(0, k, j)] = -0.5, 0.5
        for i in range(2):
            zl[wl.ix(0, k, i)], zu[wl.ix(0, k, i)] = 0.0, 7.0
    zl[wl.it0(0)] = zu[wl.it0(0)] = 0.0
    zl[wl.itf(0)] = zu[wl.itf(0)] = tspan
    wl.zl, wl.zu = zl, zu
    gl, gu = np.zeros(wl.ncons), np.zeros(wl.ncons)
    ne0 = 2 * N
    gl[ne0:ne0 + 2], gu[ne0:ne0 + 2] = x0, x0
    gl[ne0 + 2:ne0 + 4], gu[ne0 + 2:ne0 + 4] = xf - xtol, xf + xtol
    gl[ne0 + 4:ne0 + 4 + wl.npath[0] * N] = -1000.0
    gl[-1], gu[-1] = 0.0, np.inf
    wl.gl, wl.gu = np.tile(gl, (batch, 1)), np.tile(gu, (batch, 1))
    wl.meta = dict(x0=x0, xf=xf, xtol=xtol, dt=dt, nsteps=nsteps)
    # guess: straight line start -> goal, zero controls; fixed times
    guess = np.zeros((batch, wl.nvars))
    s = np.linspace(0.0, 1.0, N)
    for i in range(2):
        guess[:, [wl.ix(0, k, i) for k in range(N)]] = x0[i] + s * (xf[i] - x0[i])
    guess[:, wl.itf(0)] = tspan
    wl.x = _decision_vectors(wl, guess, rng)
    return wl


# ---- C1/C2/C5 (pm3d), C3 (fw6), C4 (multi-phase pm3d) -----------------------------------------------
def _cylinder_field(rng, batch, ncyl, start, goal, box=(1000.0, 1000.0), rad=(20.0, 60.0)):
    """ncyl cylinders per instance, centres U in the box, radii U(20,60), not containing start/goal."""
    cyl = np.zeros((batch, ncyl, 3))
    todo = np.ones((batch, ncyl), dtype=bool)
    for _ in range(64):
        n = int(todo.sum())
        if n == 0:
            break
        cx, cy = rng.uniform(0, box[0], n), rng.uniform(0, box[1], n)
        r = rng.uniform(rad[0], rad[1], n)
        bi = np.nonzero(todo)[0]
        ds = np.hypot(cx - start[bi, 0], cy - start[bi, 1])
        dg = np.hypot(cx - goal[bi, 0], cy - goal[bi, 1])
        ok = (ds > r + 5.0) & (dg > r + 5.0)
        idx = np.argwhere(todo)
        good = idx[ok]
        cyl[good[:, 0], good[:, 1], 0] = cx[ok]
        cyl[good[:, 0], good[:, 1], 1] = cy[ok]
        cyl[good[:, 0], good[:, 1], 2] = r[ok]
        todo[good[:, 0], good[:, 1]] = False
    if todo.any():
        raise RuntimeError("cylinder rejection sampling did not converge")
    return cyl


def _start_goal(rng, batch, box=(1000.0, 1000.0, 120.0), minsep=600.0):
    start = np.zeros((batch, 3))
    goal = np.zeros((batch, 3))
    todo = np.ones(batch, dtype=bool)
    for _ in range(256):
        n = int(todo.sum())
        if n == 0:
            break
        s = rng.uniform(0, 1, (n, 3)) * np.array(box)
        g = rng.uniform(0, 1, (n, 3)) * np.array(box)
        ok = np.linalg.norm(s - g, axis=1) >= minsep
        bi = np.nonzero(todo)[0][ok]
        start[bi], goal[bi] = s[ok], g[ok]
        todo[bi] = False
    if todo.any():
        raise RuntimeError("start/goal sampling did not converge")
    return start, goal


_PM3D_LO = np.array([0.0, 0.0, 0.0, -50.0, -50.0, -50.0])
_PM3D_HI = np.array([1000.0, 1000.0, 120.0, 50.0, 50.0, 50.0])
_PM3D_ULO, _PM3D_UHI = np.full(3, -10.0), np.full(3, 10.0)
_FW6_LO = np.array([0.0, 0.0, 0.0, 15.0, -0.5, -2.0 * np.pi])
_FW6_HI = np.array([1000.0, 1000.0, 120.0, 40.0, 0.5, 2.0 * np.pi])
_FW6_ULO, _FW6_UHI = np.array([-3.0, -0.3, -0.5]), np.array([3.0, 0.3, 0.5])


def _uas(name, model, batch, nnodes, ncyl, seed, tspan, scaled, **kw):
    rng = _rng(seed)
    P = len(nnodes)
    wl = Workload(name=name, model=model, nnodes=list(nnodes), nstatic=[ncyl] * P, batch=batch, **kw)
    ns, nc = wl.ns, wl.nc
    start, goal = _start_goal(rng, batch)
    wl.cylinders = _cylinder_field(rng, batch, ncyl * P, start, goal)
    xlo, xhi = (_PM3D_LO, _PM3D_HI) if model == PM3D else (_FW6_LO, _FW6_HI)
    ulo, uhi = (_PM3D_ULO, _PM3D_UHI) if model == PM3D else (_FW6_ULO, _FW6_UHI)
    zl, zu = np.zeros(wl.nvars), np.zeros(wl.nvars)
    guess = np.zeros((batch, wl.nvars))
    gl = np.zeros((batch, wl.ncons))
    gu = np.zeros((batch, wl.ncons))
    xtol = np.full(ns, 0.5)
    for p, N in enumerate(nnodes):
        ta, tb = tspan * p / P, tspan * (p + 1) / P
        s = (np.linspace(0.0, 1.0, N) + p) / P  # fraction of the whole path at each node
        for k in range(N):
            zl[wl.ix(p, k, 0):wl.ix(p, k, 0) + ns] = xlo
            zu[wl.ix(p, k, 0):wl.ix(p, k, 0) + ns] = xhi
            zl[wl.iu(p, k, 0):wl.iu(p, k, 0) + nc] = ulo
            zu[wl.iu(p, k, 0):wl.iu(p, k, 0) + nc] = uhi
            for i in range(3):
                guess[:, wl.ix(p, k, i)] = start[:, i] + s[k] * (goal[:, i] - start[:, i])
            if model == FW6:
                guess[:, wl.ix(p, k, 3)] = 25.0
                guess[:, wl.ix(p, k, 5)] = np.arctan2(goal[:, 1] - start[:, 1], goal[:, 0] - start[:, 0])
            else:
                for i in range(3):
                    guess[:, wl.ix(p, k, 3 + i)] = (goal[:, i] - start[:, i]) / tspan
        zl[wl.it0(p)] = zu[wl.it0(p)] = ta
        zl[wl.itf(p)] = zu[wl.itf(p)] = tb
        guess[:, wl.it0(p)], guess[:, wl.itf(p)] = ta, tb
        # constraint bounds: defects 0; events [x0,x0] / [xf-tol, xf+tol] (ePSOPT.cpp:137-141) at the
        # phase boundaries of the straight-line guess; path [-1000, 0]; tf - t0 >= 0
        go = wl.goff(p)
        e0 = go + ns * N
        xs = np.stack([guess[:, wl.ix(p, 0, i)] for i in range(ns)], axis=1)
        xe = np.stack([guess[:, wl.ix(p, N - 1, i)] for i in range(ns)], axis=1)
        gl[:, e0:e0 + ns], gu[:, e0:e0 + ns] = xs, xs
        gl[:, e0 + ns:e0 + 2 * ns], gu[:, e0 + ns:e0 + 2 * ns] = xe - xtol, xe + xtol
        p0 = e0 + 2 * ns
        gl[:, p0:p0 + ncyl * N] = -1000.0
        gu[:, p0 + ncyl * N] = np.inf
    wl.zl, wl.zu, wl.gl, wl.gu = zl, zu, gl, gu
    wl.meta = dict(start=start, goal=goal, tspan=tspan)
    if scaled:
        wl.sz = psopt_like_scaling(wl)
    wl.x = _decision_vectors(wl, guess, rng)
    return wl


def pm3d(batch=4096, nnodes=40, ncyl=8, seed=SEED, scaled=False, **kw):
    """C1 (batch=1), C2 (batch=4096), C5 (batch=65536): 3-D point mass, 8 cylinders, 40 LGL nodes."""
    return _uas(f"C2-pm3d-N{nnodes}-B{batch}", PM3D, batch, [nnodes], ncyl, seed, 60.0, scaled, **kw)


def fw6(batch=64, nnodes=200, ncyl=64, seed=SEED, scaled=False, **kw):
    """C3: 6-state fixed-wing kinematics, 200 nodes, 64 cylinders."""
    return _uas(f"C3-fw6-N{nnodes}-B{batch}", FW6, batch, [nnodes], ncyl, seed, 60.0, scaled, **kw)


def pm3d_multiphase(batch=1024, nphases=3, nnodes=30, ncyl=8, seed=SEED, scaled=False, **kw):
    """C4: takeoff/cruise/landing as 3 pm3d phases linked by state + time continuity."""
    return _uas(f"C4-pm3d-{nphases}x{nnodes}-B{batch}", PM3D, batch, [nnodes] * nphases, ncyl, seed, 90.0,
                scaled, **kw)


# ---- user models (dynamics and cost recorded from callbacks, etol_b200.tape) ---------------------------
def pm3d_user(batch=8, **kw):
    """pm3d written as callbacks and registered as a user model: same data as `pm3d`, runtime-compiled kernels."""
    from . import capi, tape as T
    wl = pm3d(batch=batch, **kw)
    wl.tape = T.pm3d_tape()
    wl.model = capi.register_user_model(wl.tape)
    wl.name = wl.name.replace("pm3d", "pm3d-as-user-model")
    return wl


def planar_user(tape, batch=8, nnodes=33, ncyl=5, ntracks=0, nwaypoints=3, seed=SEED, name="user", scaled=False,
                xlo=(0.0, 0.0, -8.0, -8.0), xhi=(1000.0, 1000.0, 8.0, 8.0), ulo=(-3.0, -3.0), uhi=(3.0, 3.0),
                rest=(0.3, 5.0), **kw):
    """A 4-state, 2-control planar user model (states 0, 1 = position) among cylinders and moving circles.
    `rest` is the guess of states 2 and 3; tf is free in [30, 90] s."""
    from . import capi
    rng = _rng(seed)
    wl = Workload(name=f"{name}-N{nnodes}-B{batch}", model=capi.register_user_model(tape), nnodes=[nnodes],
                  nstatic=[ncyl], batch=batch, ntracks=ntracks, nwaypoints=nwaypoints if ntracks else 0, tape=tape, **kw)
    ns, nc, N = wl.ns, wl.nc, nnodes
    assert (ns, nc) == (4, 2)
    start, goal = _start_goal(rng, batch)
    wl.cylinders = _cylinder_field(rng, batch, ncyl, start, goal)
    if ntracks:
        wl.tracks = []
        for b in range(batch):
            trk = []
            for _ in range(ntracks):
                tt = np.linspace(0.0, 90.0, nwaypoints)
                trk.append((float(rng.uniform(10.0, 30.0)), tt, rng.uniform(0.0, 1000.0, nwaypoints),
                            rng.uniform(0.0, 1000.0, nwaypoints)))
            wl.tracks.append(trk)
    zl, zu = np.zeros(wl.nvars), np.zeros(wl.nvars)
    guess = np.zeros((batch, wl.nvars))
    s = np.linspace(0.0, 1.0, N)
    for k in range(N):
        zl[wl.ix(0, k, 0):wl.ix(0, k, 0) + ns], zu[wl.ix(0, k, 0):wl.ix(0, k, 0) + ns] = xlo, xhi
        zl[wl.iu(0, k, 0):wl.iu(0, k, 0) + nc], zu[wl.iu(0, k, 0):wl.iu(0, k, 0) + nc] = ulo, uhi
        for i in range(2):
            guess[:, wl.ix(0, k, i)] = start[:, i] + s[k] * (goal[:, i] - start[:, i])
        guess[:, wl.ix(0, k, 2)], guess[:, wl.ix(0, k, 3)] = rest
    zl[wl.it0(0)] = zu[wl.it0(0)] = 0.0
    zl[wl.itf(0)], zu[wl.itf(0)] = 30.0, 90.0
    guess[:, wl.itf(0)] = 60.0
    gl, gu = np.zeros((batch, wl.ncons)), np.zeros((batch, wl.ncons))
    e0 = ns * N
    xs = np.stack([guess[:, wl.ix(0, 0, i)] for i in range(ns)], axis=1)
    xe = np.stack([guess[:, wl.ix(0, N - 1, i)] for i in range(ns)], axis=1)
    gl[:, e0:e0 + ns], gu[:, e0:e0 + ns] = xs, xs
    gl[:, e0 + ns:e0 + 2 * ns], gu[:, e0 + ns:e0 + 2 * ns] = xe - 0.5, xe + 0.5
    p0 = e0 + 2 * ns
    gl[:, p0:p0 + wl.npath[0] * N] = -1000.0
    gu[:, p0 + wl.npath[0] * N] = np.inf
    wl.zl, wl.zu, wl.gl, wl.gu = zl, zu, gl, gu
    wl.meta = dict(start=start, goal=goal)
    if scaled:
        wl.sz = psopt_like_scaling(wl)
    wl.x = _decision_vectors(wl, guess, rng)
    return wl


def unicycle(batch=8, **kw):
    from . import tape as T
    return planar_user(T.unicycle_tape(), batch=batch, name="user-unicycle", xlo=(0.0, 0.0, -3.2, 0.0),
                       xhi=(1000.0, 1000.0, 3.2, 25.0), ulo=(-2.0, -0.5), uhi=(2.0, 0.5), rest=(0.6, 12.0), **kw)


def dragmass(batch=8, **kw):
    from . import tape as T
    return planar_user(T.drag_tape(), batch=batch, name="user-dragmass", rest=(4.0, -3.0), **kw)


def gust(batch=8, **kw):
    """planar point mass with drag in a time-varying wind, time-weighted cost: dynamics and cost read t"""
    from . import tape as T
    return planar_user(T.gust_tape(), batch=batch, name="user-gust", rest=(4.0, -3.0), **kw)


def zone(batch=8, timedep=False, **kw):
    """drag model (optionally in the time-varying wind) with three traced path constraints: a growing disc, a
    drifting ellipse and a wall -- none of them a built-in zone row"""
    from . import tape as T
    return planar_user(T.gust_zone_tape() if timedep else T.zone_tape(), batch=batch, name="user-zone", rest=(4.0, -3.0), **kw)
