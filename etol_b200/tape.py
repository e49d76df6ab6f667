"""Recording scalar for user models: the Python mirror of `ecuda::var` (include/ETOL/eCUDA_var.hpp).

ePSOPT evaluates user lambdas on ADOL-C adoubles (src/ePSOPT/ePSOPT.cpp:186-276 of the ETOL tree);
eCUDA runs them once on a recording scalar and ships the resulting tape to the library
(`ecuda_register_user_model`, include/ecuda.h), which differentiates it and compiles the kernels for it.
`trace(ns, nc, dynamics, cost)` calls `dynamics(x, u) -> [dx_i/dt]` and `cost(x, u) -> L` on `Var`s.
"""
import math
from dataclasses import dataclass, field
from typing import List, Tuple

OP_INPUT, OP_CONST, OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_NEG, OP_POW, OP_SQRT, OP_SIN, OP_COS, OP_EXP = range(12)
STATIC_CYLINDER, STATIC_EDGE = 0, 1


@dataclass
class Tape:
    ns: int
    nc: int
    static_kind: int = STATIC_CYLINDER
    nodes: List[Tuple[int, int, int, float]] = field(default_factory=list)  # (op, a, b, imm)
    f_out: List[int] = field(default_factory=list)
    cost_out: int = -1
    row_out: List[int] = field(default_factory=list)  # traced path rows (may read states 0, 1 and t)

    def push(self, op, a=-1, b=-1, imm=0.0):
        self.nodes.append((op, a, b, float(imm)))
        return len(self.nodes) - 1

    def key(self):
        return (self.ns, self.nc, self.static_kind, tuple(self.nodes), tuple(self.f_out), self.cost_out, tuple(self.row_out))


class Var:
    __slots__ = ("tape", "id")

    def __init__(self, tape, id_):
        self.tape, self.id = tape, id_

    def _lift(self, other):
        if isinstance(other, Var):
            return other
        return Var(self.tape, self.tape.push(OP_CONST, imm=float(other)))

    def _bin(self, op, other, swap=False):
        o = self._lift(other)
        a, b = (o, self) if swap else (self, o)
        return Var(self.tape, self.tape.push(op, a.id, b.id))

    def __add__(self, o): return self._bin(OP_ADD, o)
    def __radd__(self, o): return self._bin(OP_ADD, o, True)
    def __sub__(self, o): return self._bin(OP_SUB, o)
    def __rsub__(self, o): return self._bin(OP_SUB, o, True)
    def __mul__(self, o): return self._bin(OP_MUL, o)
    def __rmul__(self, o): return self._bin(OP_MUL, o, True)
    def __truediv__(self, o): return self._bin(OP_DIV, o)
    def __rtruediv__(self, o): return self._bin(OP_DIV, o, True)
    def __neg__(self): return Var(self.tape, self.tape.push(OP_NEG, self.id))
    def __pow__(self, e): return Var(self.tape, self.tape.push(OP_POW, self.id, imm=float(e)))


def _un(op, v):
    return Var(v.tape, v.tape.push(op, v.id))


def sqrt(v): return _un(OP_SQRT, v)
def sin(v): return _un(OP_SIN, v)
def cos(v): return _un(OP_COS, v)
def exp(v): return _un(OP_EXP, v)


def trace(ns, nc, dynamics, cost, static_kind=STATIC_CYLINDER, with_time=False, rows=None):
    """with_time: the callbacks also receive the node time, `dynamics(x, u, t)` / `cost(x, u, t)` -- the `k` argument
    of the ePSOPT callbacks (src/ePSOPT/ePSOPT.cpp:218-260 of the ETOL tree)."""
    t = Tape(ns, nc, static_kind)
    x = [Var(t, t.push(OP_INPUT, i)) for i in range(ns)]
    u = [Var(t, t.push(OP_INPUT, ns + j)) for j in range(nc)]
    extra = (Var(t, t.push(OP_INPUT, ns + nc)),) if with_time else ()
    f = dynamics(x, u, *extra)
    assert len(f) == ns, "one state derivative per state"
    t.f_out = [x[0]._lift(v).id for v in f]
    t.cost_out = x[0]._lift(cost(x, u, *extra)).id
    if rows is not None:
        # traced path constraints: `rows(x0, x1, t) -> [values]`, evaluated at every node after the built-in zone rows
        # (what ePSOPT does with every entry of _constraints, src/ePSOPT/ePSOPT.cpp:262-270)
        tv = extra[0] if with_time else Var(t, t.push(OP_INPUT, ns + nc))
        t.row_out = [x[0]._lift(v).id for v in rows(x[0], x[1], tv)]
    return t


# ---- the user models of the test-suite / examples --------------------------------------------------------
def pm3d_tape():
    """The built-in pm3d model written as callbacks: must reproduce it bit for bit."""
    return trace(6, 3, lambda x, u: [x[3], x[4], x[5], u[0], u[1], u[2]],
                 lambda x, u: (u[0] * u[0] + u[1] * u[1]) + u[2] * u[2])


def unicycle_tape():
    """Planar unicycle with speed state: x' = v cos th, y' = v sin th, th' = w, v' = a; cost a^2 + w^2."""
    return trace(4, 2, lambda x, u: [x[3] * cos(x[2]), x[3] * sin(x[2]), u[1], u[0]],
                 lambda x, u: u[0] * u[0] + u[1] * u[1])


def drag_tape(cd=0.02):
    """Planar point mass with quadratic drag (every velocity derivative reads its own state):
    p' = v, v' = a - cd*|v|*v with |v| = sqrt(vx^2 + vy^2 + 1e-6); cost |a|^2 + 0.1*|v|^2."""
    def dyn(x, u):
        speed = sqrt(x[2] ** 2 + x[3] ** 2 + 1e-6)
        return [x[2], x[3], u[0] - cd * speed * x[2], u[1] - cd * speed * x[3]]
    return trace(4, 2, dyn, lambda x, u: (u[0] * u[0] + u[1] * u[1]) + 0.1 * (x[2] * x[2] + x[3] * x[3]))


def gust_tape(cd=0.02, w0=1.5, omega=0.11):
    """Time-dependent user model: the drag model in a wind that turns and breathes with time, and a running cost
    whose weight grows along the flight. p' = v + w(t), v' = a - cd*|v|*v with w(t) = w0*(1 + 0.02 t)*(cos, sin)(omega t);
    cost (|a|^2 + 0.1*|v|^2)*(1 + 0.01 t). Both the dynamics and the cost read t."""
    def dyn(x, u, t):
        speed = sqrt(x[2] ** 2 + x[3] ** 2 + 1e-6)
        amp = w0 * (1.0 + 0.02 * t)
        return [x[2] + amp * cos(omega * t), x[3] + amp * sin(omega * t),
                u[0] - cd * speed * x[2], u[1] - cd * speed * x[3]]
    return trace(4, 2, dyn, lambda x, u, t: ((u[0] * u[0] + u[1] * u[1]) + 0.1 * (x[2] * x[2] + x[3] * x[3])) * (1.0 + 0.01 * t),
                 with_time=True)


def zone_rows(x0, x1, t):
    """Path constraints none of the built-in zone rows can express, written as callbacks: a keep-out disc that grows
    with time, an ellipse whose centre drifts with time, and a wall that only reads the first position state."""
    grow = 40.0 + 0.8 * t
    disc = grow * grow - ((x0 - 500.0) ** 2 + (x1 - 500.0) ** 2)
    cx, cy = 200.0 + 4.0 * t, 700.0 - 3.0 * t
    ell = 1.0 - (((x0 - cx) / 60.0) ** 2 + ((x1 - cy) / 30.0) ** 2)
    wall = (x0 - 990.0) * 0.5
    return [disc, ell, wall]


def zone_tape(cd=0.02):
    """the drag model with three traced path rows (zone_rows)"""
    def dyn(x, u):
        speed = sqrt(x[2] ** 2 + x[3] ** 2 + 1e-6)
        return [x[2], x[3], u[0] - cd * speed * x[2], u[1] - cd * speed * x[3]]
    return trace(4, 2, dyn, lambda x, u: (u[0] * u[0] + u[1] * u[1]) + 0.1 * (x[2] * x[2] + x[3] * x[3]), rows=zone_rows)


def gust_zone_tape():
    """time-dependent dynamics and cost (gust_tape) together with the traced path rows"""
    g = gust_tape()
    def dyn(x, u, t):
        speed = sqrt(x[2] ** 2 + x[3] ** 2 + 1e-6)
        amp = 1.5 * (1.0 + 0.02 * t)
        return [x[2] + amp * cos(0.11 * t), x[3] + amp * sin(0.11 * t), u[0] - 0.02 * speed * x[2], u[1] - 0.02 * speed * x[3]]
    return trace(4, 2, dyn, lambda x, u, t: ((u[0] * u[0] + u[1] * u[1]) + 0.1 * (x[2] * x[2] + x[3] * x[3])) * (1.0 + 0.01 * t),
                 with_time=True, rows=zone_rows)
