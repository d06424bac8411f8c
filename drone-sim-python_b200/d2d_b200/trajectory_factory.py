"""Named trajectories -- the `d2d.trajectory_factory` registry (d2d/trajectory_factory.py) for the
trajectory families the engine evaluates, spline, tabulated and space-indexed ones included."""
import numpy as np

from . import _lib
from . import trajectory as ddt

trajectories = {}


def register(T):
    trajectories[T.name] = (T.desc, T)
    return T


def list_available():
    return [f"{k}: {v[0]}" for k, v in sorted(trajectories.items())]


@register
class TrajCircle(ddt.TrajectoryCircle):                      # d2d/trajectory_factory.py:20-25
    name, desc = "circle", "30m radius (30,30) centered circle"
    extends = (-10, 70, -10, 70)

    def __init__(self):
        super().__init__(c=[30., 30.], r=30., v=10., t0=0., alpha0=0, dalpha=2 * np.pi)


@register
class TrajTwoLines(ddt.CompositeTraj):                       # :29-37
    name, desc = "two_lines", "example of composite trajectory"
    extends = (-20, 120, -20, 60)

    def __init__(self):
        first = ddt.TrajectoryLine([0, 0], [50, 50], v=10., t0=0.)
        super().__init__([first, ddt.TrajectoryLine([50, 50], [100, 0], v=10., t0=first.duration)])


@register
class TrajSquare(ddt.CompositeTraj):                         # :40-50
    name, desc = "square", "example of composite trajectory"
    extends = (-10, 60, -10, 60)

    def __init__(self):
        corners = [[0, 0], [50, 0], [50, 50], [0, 50]]
        sides, t0 = [], 0.
        for k in range(4):
            sides.append(ddt.TrajectoryLine(corners[k], corners[(k + 1) % 4], v=10., t0=t0))
            t0 = t0 + sides[-1].duration
        super().__init__(sides)


@register
class TrajLineWithIntro(ddt.CompositeTraj):                  # :53-61
    name, desc = "line_with_intro", "line with circle_intro"

    def __init__(self, Y0=[0, 0], Y1=[0, 50], Y2=[100, 50], r=-25.):
        arc = ddt.TrajectoryCircle(c=(np.asarray(Y0) + Y1) / 2, r=r, v=10., alpha0=np.pi / 2, dalpha=np.pi)
        super().__init__([arc, ddt.TrajectoryLine(Y1, Y2, v=10., t0=arc.duration)])
        self.extends = (-50, 130, -30, 130)


class TrajWithIntro(ddt.CompositeTraj):                      # :65-87 (not registered there either)
    name, desc = "traj_with_intro", "traj with circle_intro"

    def __init__(self, Y0, traj, v=10, duration=8.):
        start = traj.get(0)[0]
        gap = np.linalg.norm(start - Y0)
        super().__init__([ddt.TrajectoryLine(Y0, start, v=gap / duration), traj])


@register
class TrajMinSnapDemo(ddt.MinSnapPoly):                      # :110-117
    name, desc = "demo_minsnap", "demo_minsnap"
    extends = (-10, 210, -10, 210)

    def __init__(self):
        super().__init__([[0, 10, 0, 0], [0, 0, 0, 0]], [[200, 0, 0, 0], [200, 10, 0, 0]], duration=33.65)


@register
class TrajSlalom(ddt.Trajectory):                            # :121-145 (a = 10, om = 1 fixed at :138)
    name, desc = "slalom", "slalom"
    extends = (-10, 100, -10, 50)

    def __init__(self, p1=[0, 20], p2=[100, 20], v=10., t0=0., phi=0.):
        self.p1, self.p2, self.v, self.t0, self.phi = np.asarray(p1), np.asarray(p2), v, t0, phi
        dep = self.p2 - self.p1
        self.length = np.linalg.norm(dep)
        self.un = dep / self.length
        self.duration = self.length / self.v

    def segments(self):
        uv = self.un * self.v
        p = np.zeros(_lib.SEG_NPAR)
        p[0], p[1], p[2], p[3], p[4], p[5] = self.t0, self.p1[0], self.p1[1], uv[0], uv[1], self.phi
        return [(_lib.SEG_SLALOM, p)]


@register
class TrajSiDemo(ddt.SpaceIndexedTraj):                      # :177-185
    name, desc = "sidemo", "space indexed trajectory demo"
    extends = (-10, 100, -10, 50)

    def __init__(self, p1=[0, 20], p2=[100, 20], duration=10., t0=0.):
        dynamic = ddt.PolynomialOne([0, 0.05, 0, 0], [1, 0.05, 0, 0], duration=duration)
        super().__init__(ddt.TrajectoryLine([0, 20], [100, 20], v=100), dynamic)


@register
class TrajTabulated(ddt.Trajectory):                         # d2d/trajectory_factory.py:149-171
    """Zero-order lookup of a planner solution (.npz with the keys of Planner.save_solution, 06_optyplan.py:135-139):
    position of the first stored sample at or after t, ground velocity v (cos psi, sin psi) + wind, no higher derivatives."""
    name, desc = "tabulated", "tabulated"
    extends = (-5, 25, -10, 20)

    def __init__(self, filename="./optyplan_exp0.npz"):
        d = np.load(filename)
        self.sol_time, self.sol_x, self.sol_y, self.sol_psi, self.sol_phi, self.sol_v, self.wind = \
            [d[k] for k in ("sol_time", "sol_x", "sol_y", "sol_psi", "sol_phi", "sol_v", "wind")]
        self.t0 = 0.
        self.duration = self.sol_time[-1]

    def table(self):
        vx = self.sol_v * np.cos(self.sol_psi) + self.wind[:, 0]
        vy = self.sol_v * np.sin(self.sol_psi) + self.wind[:, 1]
        return np.stack([self.sol_time, self.sol_x, self.sol_y, vx, vy]).astype(np.float64)

    def segments(self):
        return [(_lib.SEG_TABLE, np.zeros(_lib.SEG_NPAR))]


def _spline_pieces(xs, ys, k=4):
    """Quartic interpolating spline (FITPACK through scipy, as upstream: host-side setup) -> break points and
    per-interval polynomial coefficients in ascending powers of (t - break point), padded to 8."""
    import scipy.interpolate as interpolate
    spl = interpolate.InterpolatedUnivariateSpline(xs, ys, k=k)
    pp = interpolate.PPoly.from_spline(spl._eval_args)
    keep = np.nonzero(np.diff(pp.x) > 0)[0]                   # repeated end knots give empty intervals
    coefs = np.zeros((len(keep), 8))
    coefs[:, :k + 1] = pp.c[::-1, keep].T
    return pp.x[keep], pp.x[keep + 1], coefs, spl


@register
class TrajSpline(ddt.Trajectory):                            # d2d/trajectory_factory.py:189-211
    """Quartic spline through way points, periodic in time.  The spline's polynomial pieces become SEG_POLY segments of a
    composite trajectory, so `get` and the rollouts evaluate them on the device (FITPACK's de Boor recursion and the
    piecewise polynomial agree to ~1e-13)."""
    name, desc = "spline", "spline dev"

    def __init__(self, waypoints=None, duration=None):
        self.waypoints = np.array([[0., 0.], [50, 50], [100, 0], [150, 50], [200, 0]]) if waypoints is None else np.asarray(waypoints, dtype=float)
        if duration is None:
            v = 10.
            duration = np.sum(np.linalg.norm(self.waypoints[1:] - self.waypoints[:-1], axis=1)) / v
        self.duration = duration
        lam = np.linspace(0, self.duration, len(self.waypoints))
        x0, x1, cx, sx = _spline_pieces(lam, self.waypoints[:, 0])
        _, _, cy, sy = _spline_pieces(lam, self.waypoints[:, 1])
        self.splines = [sx, sy]
        self._starts, self.steps_end, self._cx, self._cy = x0, x1, cx, cy
        self.extends = [0, 200, -20, 80]
        self.t0 = 0.

    def is_composite(self):
        return True

    def segments(self):
        segs = []
        for t_i, cx, cy in zip(self._starts, self._cx, self._cy):
            p = np.zeros(_lib.SEG_NPAR)
            p[0], p[1:9], p[9:17] = t_i, cx, cy
            segs.append((_lib.SEG_POLY, p))
        return segs


class SplineOne:                                             # d2d/trajectory_factory.py:213-221
    """Scalar quartic spline; `get(t)` -> value and three derivatives, evaluated on the device."""

    def __init__(self, xs, ys):
        self.nder = 3
        self.duration = xs[-1]
        self._starts, self._ends, self._c, self.dyn = _spline_pieces(np.asarray(xs, float), np.asarray(ys, float))

    def get(self, t):
        from .engine import PackedTrajectories, get_engine
        eng = get_engine()
        n = len(self._starts)
        par = np.zeros((_lib.SEG_NPAR, n))
        par[0], par[1:9] = self._starts, self._c.T
        ends = self._ends.copy(); ends[-1] = np.inf             # no wrap: the last piece extends (upstream extrapolates too)
        tab = eng.table(PackedTrajectories([0], [n], [0.], [np.inf], [_lib.SEG_POLY] * n, ends, par))
        Y = eng.traj_eval(tab, eng.to_device(np.array([float(t)])))
        return Y.cpu().numpy().reshape(4, 2)[:, 0].copy()


class FooOne:                                                # d2d/trajectory_factory.py:225-234
    """Piecewise-linear scalar dynamic through (xs, ys): value, slope, 0, 0."""

    def __init__(self, xs, ys):
        self.xs, self.ys = np.asarray(xs), np.asarray(ys)
        self.dxs, self.dys = self.xs[1:] - self.xs[:-1], self.ys[1:] - self.ys[:-1]
        self.ds = self.dys / self.dxs
        self.duration = xs[-1]

    def get(self, t):
        _i = np.where(t >= self.xs)[0][-1]
        return np.array([self.ys[_i] + (t - self.xs[_i]) * self.ds[_i], self.ds[_i], 0, 0])


@register
class TrajSiSpline(ddt.SpaceIndexedTraj):                    # d2d/trajectory_factory.py:241-285
    """Three quarters of a circle flown at (nearly) constant AIR speed in a 5 m/s wind: the space-indexed circle whose
    piecewise-linear dynamic lambda(t) the upstream constructor finds with scipy.optimize.minimize (:264-279).  Like upstream,
    the object is left with the LAST dynamic the optimiser probed (err_fun calls set_dyn; the fitted spline goes to `_dyn4`
    and is not installed).  Every evaluation of the flat output inside the optimisation runs on the engine (one traj_eval
    launch per probe); the linear pieces become space-indexed segments of one composite trajectory.

    `knots=(xs, ys)` installs a given dynamic instead of running the optimiser (a BFGS path is not reproducible to the last
    bit across BLAS / libm builds: parity tests feed the knots the unmodified reference ended with)."""
    name, desc = "sispline", "spline dev"
    extends = (-10, 100, -10, 50)

    def __init__(self, duration=30., knots=None, vtarget=10., wind=(5., 0.)):
        geometry = ddt.TrajectoryCircle(c=[30., 30.], r=30., v=2 * np.pi * 30., t0=0., alpha0=0, dalpha=3 * np.pi / 2)
        dynamic = ddt.AffineOne(1. / duration, 0., duration=duration)
        ddt.SpaceIndexedTraj.__init__(self, geometry, dynamic)
        self._dyn1 = self._dyn
        self.ts = np.arange(0, self._dyn.duration, 0.5)
        if knots is not None:
            self.set_dyn(FooOne(np.asarray(knots[0], dtype=float), np.asarray(knots[1], dtype=float)))
            self._dyn4 = SplineOne(self._dyn.xs, self._dyn.ys)
            return
        npts = 10
        xs = np.linspace(0, 30, npts)
        w = np.asarray(wind, dtype=float)

        def err_fun(p):
            ys = np.concatenate(([0.], np.cumsum(p)))
            self.set_dyn(FooOne(xs, ys))
            flat_out = self.get_many(self.ts)
            vair = np.linalg.norm(flat_out[:, 1] - w, axis=1)
            return np.mean(np.square(vair - vtarget))
        import scipy.optimize
        res = scipy.optimize.minimize(err_fun, [1. / npts] * (npts - 1))
        self.fit_error = float(res.fun)
        ys = np.concatenate(([0.], np.cumsum(res.x)))
        self._dyn4 = SplineOne(xs, ys)


def print_available():
    print("Available trajectories:")
    for i, n in enumerate(list_available()):
        print(f"{i} -> {n}")


def get(traj_name, *args):
    return trajectories[traj_name][1](*args), trajectories[traj_name][0]
