"""2D+t reference trajectories -- the `d2d.trajectory` call surface (d2d/trajectory.py) as parameter
holders for the engine.  `get(t)` keeps the reference's contract (a (4,2) array: row k = k-th time
derivative of (x, y)) but is evaluated by the CUDA trajectory kernel; rollouts never call it, they
consume the packed segment table (`pack`)."""
import numpy as np

from . import _lib
from .engine import PackedTrajectories, get_engine


class Trajectory:
    """Base class (d2d/trajectory.py:88-122)."""
    desc = ""
    cx, cy, ncomp = 0, 1, 2
    nder = 3
    extends = (0, 100, 0, 100)

    def __init__(self):
        self.t0 = 0.

    def reset(self, t0):
        self.t0 = t0

    # -- engine plumbing --------------------------------------------------------------------------
    def segments(self):
        """[(seg_type, par[SEG_NPAR])] in evaluation order; composites override."""
        raise NotImplementedError

    def is_composite(self):
        return False

    def get_many(self, ts, engine=None):
        """Trajectory.get for an array of times -> (len(ts), 4, 2)."""
        eng = engine or get_engine()
        ts = np.atleast_1d(np.asarray(ts, dtype=np.float64))
        tab = eng.table(pack([self]))
        Y = eng.traj_eval(tab, eng.to_device(ts))            # [nT][8][1]
        return Y.cpu().numpy().reshape(len(ts), 4, 2)

    def get(self, t):
        return self.get_many([t])[0]

    def compute_extends(self, dt=0.1):                        # d2d/trajectory.py:102-115
        Ys = self.get_many(np.arange(self.t0, self.t0 + self.duration, dt))
        p0, p1 = np.min(Ys[:, 0], axis=0).round(1) - 1, np.max(Ys[:, 0], axis=0).round(1) + 1
        self.extends = (p0[0], p1[0], p0[1], p1[1])

    def summarize(self):
        return f"{self.desc}\nduration: {self.duration:.2f}s\nextends: {self.extends}"


def _par(**slots):
    p = np.zeros(_lib.SEG_NPAR)
    for k, v in slots.items():
        p[int(k[1:])] = v
    return p


class TrajectoryLine(Trajectory):
    """Constant-velocity straight line from p1 towards p2 (d2d/trajectory.py:125-141)."""

    def __init__(self, p1, p2, v=10., t0=0.):
        self.p1, self.p2, self.v, self.t0 = np.asarray(p1), np.asarray(p2), v, t0
        dep = self.p2 - self.p1
        self.length = np.linalg.norm(dep)
        self.un = dep / self.length
        self.duration = self.length / self.v

    def segments(self):
        uv = self.un * self.v
        return [(_lib.SEG_LINE, _par(s0=self.t0, s1=self.p1[0], s2=self.p1[1], s3=uv[0], s4=uv[1]))]


class TrajectoryCircle(Trajectory):
    """Circle at constant ground speed; the sign of r sets the direction (d2d/trajectory.py:143-160)."""

    def __init__(self, c=[30., 30.], r=30., v=10., t0=0., alpha0=0, dalpha=2 * np.pi):
        self.c, self.r, self.v, self.t0 = np.asarray(c), r, v, t0
        self.alpha0, self.dalpha = alpha0, dalpha
        self.omega = self.v / self.r
        self.duration = np.abs(r) * dalpha / v

    def segments(self):
        return [(_lib.SEG_CIRCLE, _par(s0=self.t0, s1=self.c[0], s2=self.c[1], s3=self.r, s4=self.omega, s5=self.alpha0))]


def arr(k, n):
    """Arrangements n!/(n-k)! (d2d/trajectory.py:41-45)."""
    a = 1
    for i in range(n, n - k, -1):
        a *= i
    return a


class CstOne:                                                # d2d/trajectory.py:13-17
    """Constant scalar trajectory; as a SpaceIndexedTraj dynamic it is a degree-0 polynomial."""

    def __init__(self, c=-1., duration=1.):
        self.c, self.duration = c, duration
        self.coefs = np.zeros((4, 8)); self.coefs[0, 0] = c

    def get(self, t):
        return np.array([self.c, 0, 0, 0])


class AffineOne:                                             # d2d/trajectory.py:19-24
    def __init__(self, c1=-1., c2=0, duration=1.):
        self.c1, self.c2, self.duration = c1, c2, duration
        self.coefs = np.zeros((4, 8)); self.coefs[0, 0], self.coefs[0, 1] = c2, c1

    def get(self, t):
        return np.array([self.c1 * t + self.c2, self.c1, 0, 0])


class SinOne:                                                # d2d/trajectory.py:26-38
    def __init__(self, c=0., a=1., om=1., duration=2 * np.pi):
        self.duration = duration
        self.c, self.a, self.om = c, a, om
        self.t0 = 0.

    def get(self, t):
        alpha = self.om * (t - self.t0)
        asa, aca = self.a * np.sin(alpha), self.a * np.cos(alpha)
        return np.array([self.c + asa, self.om * aca, -self.om ** 2 * asa, -self.om ** 3 * aca])


class PolynomialOne:
    """Scalar min-snap polynomial through boundary values and derivatives (d2d/trajectory.py:47-82).
    Coefficients are solved on the host exactly as the reference does (setup, not hot path); evaluation
    (`get`) runs on the device through a one-component space-indexed segment."""

    def __init__(self, Y0, Y1, duration):
        self.duration = duration
        nd = len(Y0)
        self._der, self._order = nd, 2 * nd
        if nd != 4:
            raise NotImplementedError("the engine evaluates 4-derivative (8-coefficient) polynomials, as every reference trajectory uses")
        self.coefs = np.zeros((nd, 2 * nd))
        M1 = np.diag([float(arr(i, i)) for i in range(nd)])
        self.coefs[0, :nd] = np.dot(np.linalg.inv(M1), Y0)
        M3, M4 = np.zeros((nd, nd)), np.zeros((nd, nd))
        for i in range(nd):
            for j in range(nd):
                if j >= i:
                    M3[i, j] = arr(i, j) * duration ** (j - i)
                M4[i, j] = arr(i, j + nd) * duration ** (j - i + nd)
        self.coefs[0, nd:] = np.dot(np.linalg.inv(M4), Y1 - np.dot(M3, self.coefs[0, :nd]))
        for d in range(1, nd):
            for pw in range(2 * nd - d):
                self.coefs[d, pw] = arr(d, pw + d) * self.coefs[0, pw + d]

    def get(self, t):
        """Value and three derivatives at t -> (4,).  Evaluated on the device as the x component of a plain polynomial
        segment whose y coefficients are zero."""
        eng = get_engine()
        par = np.zeros(_lib.SEG_NPAR)
        par[1:9] = self.coefs[0]
        tab = eng.table(PackedTrajectories([0], [1], [0.], [0.], [_lib.SEG_POLY], [0.], par.reshape(-1, 1), _lib.SEG_POLY))
        Y = eng.traj_eval(tab, eng.to_device(np.array([float(t)])))
        return Y.cpu().numpy().reshape(4, 2)[:, 0].copy()


class MinSnapPoly(Trajectory):
    """Two PolynomialOne components (d2d/trajectory.py:166-187)."""

    def __init__(self, Y00=[0, 0], Y10=[1, 0], duration=1.):
        self.duration = duration
        Y0 = np.zeros((self.ncomp, self.nder + 1))
        if np.asarray(Y00).ndim == 1: Y0[:, 0] = Y00
        else: Y0 = np.asarray(Y00, dtype=float)
        Y1 = np.zeros((self.ncomp, self.nder + 1))
        if np.asarray(Y10).ndim == 1: Y1[:, 0] = Y10
        else: Y1 = np.asarray(Y10, dtype=float)
        self._polys = [PolynomialOne(Y0[i], Y1[i], duration) for i in range(self.ncomp)]
        self.t0 = 0

    def segments(self):
        p = np.zeros(_lib.SEG_NPAR)
        p[0] = self.t0
        p[1:9] = self._polys[0].coefs[0]
        p[9:17] = self._polys[1].coefs[0]
        return [(_lib.SEG_POLY, p)]


class CompositeTraj(Trajectory):
    """Sequence of trajectories; time wraps with fmod over the total duration (d2d/trajectory.py:190-208)."""

    def __init__(self, steps):
        self.steps = steps
        self.steps_dur = [s.duration for s in self.steps]
        self.steps_end = np.cumsum(self.steps_dur)
        self.duration = np.sum(self.steps_dur)
        for s, st in zip(self.steps[1:], self.steps_end):
            s.reset(st)
        self.t0 = 0.

    def is_composite(self):
        return True

    def segments(self):
        segs = []
        for s in self.steps:
            if s.is_composite():
                raise NotImplementedError("nested CompositeTraj is not used by the reference and not supported")
            segs += s.segments()
        return segs


class TabulatedTraj(Trajectory):
    """d2d/trajectory.py:212-217: an unfinished stub upstream (the constructor ignores its file name, `get` returns zeros), kept
    for import compatibility.  The working tabulated reference is `trajectory_factory.TrajTabulated`."""

    def __init__(self, filename='/tmp/optyplan.npz'):
        pass

    def get(self, t):
        return np.zeros((self.nder + 1, self.ncomp))


class SpaceIndexedTraj(Trajectory):
    """Geometry g(lambda) driven by a scalar dynamic lambda(t) (d2d/trajectory.py:220-241).  On the engine: line or
    circle geometry with polynomial (PolynomialOne, AffineOne, CstOne) or sinusoidal (SinOne) dynamics -- TrajSiDemo
    (d2d/trajectory_factory.py:177-185) and the first stage of TrajSiSpline (:246-247)."""

    def __init__(self, geometry, dynamic):
        self.duration = dynamic.duration
        self.extends = geometry.extends
        self._geom, self._dyn = geometry, dynamic
        self.t0 = 0.

    def set_dyn(self, dyn):
        self._dyn = dyn
        self.duration = dyn.duration

    def _piecewise(self):
        """knots / values of a piecewise-linear dynamic (FooOne, d2d/trajectory_factory.py:225-234), else None"""
        dyn = self._dyn
        return (dyn.xs, dyn.ys, dyn.ds) if all(hasattr(dyn, k) for k in ("xs", "ys", "ds")) else None

    def is_composite(self):
        return self._piecewise() is not None

    @property
    def steps_end(self):
        return self._piecewise()[0][1:]

    def segments(self):
        pw = self._piecewise()
        if pw is not None:
            # one space-indexed segment per linear piece: lambda = ys[i] + (t - xs[i]) ds[i] (bit-identical: the Horner
            # recurrence in (t - knot) does the same multiply and add), selected like CompositeTraj selects its steps
            xs, ys, ds = pw
            out = []
            for i in range(len(ds)):
                ty, p = self._segment_for(_AffinePiece(ys[i], ds[i]))
                p[14] = xs[i]
                out.append((ty, p))
            return out
        return [self._segment_for(self._dyn)]

    def _segment_for(self, dyn):
        g = self._geom
        p = np.zeros(_lib.SEG_NPAR)
        if isinstance(dyn, SinOne):
            p[5], p[6], p[7], p[8], p[16] = dyn.c, dyn.a, dyn.om, dyn.t0, 1.
        elif hasattr(dyn, "coefs") and np.shape(dyn.coefs) == (4, 8):
            p[5:13] = dyn.coefs[0]
        else:
            raise NotImplementedError(f"SpaceIndexedTraj: {type(dyn).__name__} dynamics do not run on the engine "
                                      "(polynomial or SinOne dynamics do)")
        if isinstance(g, TrajectoryLine):
            uv = g.un * g.v
            p1 = np.asarray(g.p1, dtype=float) - uv * g.t0          # geometry evaluated at lambda: p1 + un v (lambda - t0)
            p[1], p[2], p[3], p[4] = p1[0], p1[1], uv[0], uv[1]
            return (_lib.SEG_SI_LINE, p)
        if isinstance(g, TrajectoryCircle):
            p[0], p[1], p[2], p[3], p[4], p[13] = g.t0, g.c[0], g.c[1], g.r, g.omega, g.alpha0
            return (_lib.SEG_SI_CIRCLE, p)
        raise NotImplementedError("SpaceIndexedTraj: only line and circle geometries run on the engine")


class _AffinePiece:
    """coefficient rows of value + slope * t in the PolynomialOne layout (4, 8)"""

    def __init__(self, value, slope):
        self.coefs = np.zeros((4, 8))
        self.coefs[0, 0], self.coefs[0, 1], self.coefs[1, 0] = value, slope, slope


class CircleBatch:
    """B plain TrajectoryCircle trajectories given as arrays (Monte-Carlo sweeps): packs without creating
    B Python objects.  Same parameters and derived quantities as TrajectoryCircle."""

    def __init__(self, cx, cy, r, v, alpha0, t0=0.):
        self.cx, self.cy, self.r, self.v, self.alpha0 = (np.asarray(a, dtype=np.float64) for a in (cx, cy, r, v, alpha0))
        self.t0 = np.broadcast_to(np.asarray(t0, dtype=np.float64), self.cx.shape)
        self.omega = self.v / self.r

    def __len__(self):
        return len(self.cx)

    def pack(self):
        B = len(self)
        par = np.zeros((_lib.SEG_NPAR, B))
        par[0], par[1], par[2], par[3], par[4], par[5] = self.t0, self.cx, self.cy, self.r, self.omega, self.alpha0
        ar = np.arange(B, dtype=np.int32)
        return PackedTrajectories(ar, np.ones(B, np.int32), np.zeros(B), np.zeros(B), np.full(B, _lib.SEG_CIRCLE, np.int32),
                                  np.zeros(B), par, _lib.SEG_CIRCLE)


class MinSnapBatch:
    """B plain MinSnapPoly trajectories given by their coefficient rows (B, 2, 8)."""

    def __init__(self, coefs0, t0=0.):
        self.coefs0 = np.asarray(coefs0, dtype=np.float64)
        self.t0 = np.broadcast_to(np.asarray(t0, dtype=np.float64), (len(self.coefs0),))

    def __len__(self):
        return len(self.coefs0)

    @classmethod
    def from_boundaries(cls, Y0, Y1, duration, t0=0.):
        """B min-snap trajectories of one common duration from boundary values Y0, Y1 of shape (B, 2, 4)
        (position and three derivatives per component) -- the algebra of PolynomialOne (d2d/trajectory.py:53-68),
        vectorised over the batch."""
        Y0, Y1 = np.asarray(Y0, dtype=np.float64), np.asarray(Y1, dtype=np.float64)
        nd = 4
        M1 = np.diag([float(arr(i, i)) for i in range(nd)])
        M3, M4 = np.zeros((nd, nd)), np.zeros((nd, nd))
        for i in range(nd):
            for j in range(nd):
                if j >= i:
                    M3[i, j] = arr(i, j) * duration ** (j - i)
                M4[i, j] = arr(i, j + nd) * duration ** (j - i + nd)
        lo = Y0 @ np.linalg.inv(M1).T
        hi = (Y1 - lo @ M3.T) @ np.linalg.inv(M4).T
        obj = cls(np.concatenate([lo, hi], axis=2), t0)
        obj.duration = duration
        return obj

    def pack(self):
        B = len(self)
        par = np.zeros((_lib.SEG_NPAR, B))
        par[0] = self.t0
        par[1:9] = self.coefs0[:, 0, :].T
        par[9:17] = self.coefs0[:, 1, :].T
        ar = np.arange(B, dtype=np.int32)
        return PackedTrajectories(ar, np.ones(B, np.int32), np.zeros(B), np.zeros(B), np.full(B, _lib.SEG_POLY, np.int32),
                                  np.zeros(B), par, _lib.SEG_POLY)


def pack(trajs):
    """Trajectory objects (or one *Batch) -> PackedTrajectories (the d2dx_traj_table layout)."""
    if hasattr(trajs, "pack"):
        return trajs.pack()
    first, nseg, t0s, durs, types, ends, pars = [], [], [], [], [], [], []
    tables, n_rows = [], 0
    for tr in trajs:
        segs = tr.segments()
        for ty, p in segs:
            if ty == _lib.SEG_TABLE:                     # p[1] is filled with the trajectory's first row in the shared tables
                tab = tr.table()
                p[1], p[2] = n_rows, tab.shape[1]
                tables.append(tab); n_rows += tab.shape[1]
        first.append(len(types)); nseg.append(len(segs))
        if tr.is_composite():
            t0s.append(float(tr.t0)); durs.append(float(tr.duration))
            ends += [float(e) for e in tr.steps_end]
        else:
            t0s.append(0.); durs.append(0.)
            ends += [0.] * len(segs)
        for ty, p in segs:
            types.append(ty); pars.append(p)
    par = np.stack(pars, axis=1)
    plain = all(d == 0. for d in durs) and all(n == 1 for n in nseg)
    uniform = types[0] if plain and len(set(types)) == 1 else -1
    return PackedTrajectories(first, nseg, t0s, durs, types, ends, par, uniform,
                              tables=np.concatenate(tables, axis=1) if tables else None)
