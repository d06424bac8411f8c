"""CPU oracle for the d2d hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain NumPy / pure-Python restatement of the reference's algorithm for the
closed-loop rollout (path A), the circular-formation rollout (path A') and the
collocation residual / Jacobian / cost / gradient evaluation (path B).  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl
reference` legs may import it, and only as the checker or the timed CPU arm.
The product package (`drone-sim-python_b200/d2d_b200`) never imports it.

Pinning: every function below is checked in `tests/test_oracle.py` against
golden vectors produced by running the UNMODIFIED reference in the build
container (`tests/golden/make_golden.py`, fixtures `tests/golden/*.npz`),
including the reference's own fixture `src/states_over_time.csv` (reproduced by
the reference to 2.9e-11 and carried in `formation.npz`) and the cached IPOPT
solutions under `src/cache/` (backward-Euler defects <= 4e-6).  What stays
"parity unpinned": the ORDER of opty's constraint / Jacobian vectors (opty is an
un-vendored, unpinned third-party dependency -- github csu-hmc/opty, imported at
src/06_optyplan.py:19 -- and is not installed here).  The values are pinned
through the reference's own sympy EoM; the ordering follows opty's published
layout and is exposed through `colloc_structure`.

All file:line citations are relative to the reference checkout's `src/`.
"""
import math

import numpy as np
import scipy.linalg

G = 9.81                      # d2d/dynamic.py:9
TAU_PHI, TAU_V = 0.01, 1.0    # d2d/dynamic.py:11-12


def norm_mpi_pi(v):
    """d2d/utils.py:7 -- floored modulo (Python / NumPy `%`)."""
    return (v + np.pi) % (2 * np.pi) - np.pi


# ---------------------------------------------------------------------------
# trajectories  (d2d/trajectory.py, d2d/trajectory_factory.py)
# ---------------------------------------------------------------------------
class Line:
    """TrajectoryLine, d2d/trajectory.py:125-141."""
    def __init__(self, p1, p2, v=10., t0=0.):
        self.p1, self.p2, self.v, self.t0 = np.asarray(p1, float), np.asarray(p2, float), v, t0
        dep = self.p2 - self.p1
        self.length = np.linalg.norm(dep)
        self.un = dep / self.length
        self.duration = self.length / self.v

    def reset(self, t0): self.t0 = t0

    def get(self, t):
        Y = np.zeros((4, 2))
        Y[0] = self.p1 + self.un * self.v * (t - self.t0)
        Y[1] = self.un * self.v
        return Y


class Circle:
    """TrajectoryCircle, d2d/trajectory.py:143-160 (sign of r gives the direction)."""
    def __init__(self, c=(30., 30.), r=30., v=10., t0=0., alpha0=0., dalpha=2 * np.pi):
        self.c, self.r, self.v, self.t0 = np.asarray(c, float), r, v, t0
        self.alpha0, self.dalpha = alpha0, dalpha
        self.omega = self.v / self.r
        self.duration = np.abs(r) * dalpha / v

    def reset(self, t0): self.t0 = t0

    def get(self, t):
        alpha = (t - self.t0) * self.omega + self.alpha0
        ca, sa = np.cos(alpha), np.sin(alpha)
        p = self.c + self.r * np.array([ca, sa])
        p1 = self.omega * self.r * np.array([-sa, ca])
        p2 = self.omega ** 2 * self.r * np.array([-ca, -sa])
        p3 = self.omega ** 3 * self.r * np.array([sa, -ca])
        return np.array((p, p1, p2, p3))


class Slalom:
    """TrajSlalom, d2d/trajectory_factory.py:121-145 (a = 10, om = 1 hard-coded at :138).
    It has no reset(): inside a CompositeTraj its t0 stays what the constructor got."""
    def __init__(self, p1=(0, 20), p2=(100, 20), v=10., t0=0., phi=0.):
        self.p1, self.p2, self.v, self.t0, self.phi = np.asarray(p1, float), np.asarray(p2, float), v, t0, phi
        dep = self.p2 - self.p1
        self.length = np.linalg.norm(dep)
        self.un = dep / self.length
        self.duration = self.length / self.v

    def reset(self, t0): self.t0 = t0     # base-class Trajectory.reset, d2d/trajectory.py:99-100

    def get(self, t):
        Y = np.zeros((4, 2))
        Y[0] = self.p1 + self.un * self.v * (t - self.t0)
        Y[1] = self.un * self.v
        a, om = 10., 1.
        alpha = om * (t - self.t0 + self.phi)
        s, c = np.sin(alpha), np.cos(alpha)
        Y[0, 1] += a * s
        Y[1, 1] += a * om * c
        Y[2, 1] += -a * om ** 2 * s
        Y[3, 1] += -a * om ** 3 * c
        return Y


def _arr(k, n):
    """arrangements n!/(n-k)!, d2d/trajectory.py:41-45."""
    a, i = 1, n
    while i > n - k:
        a *= i
        i -= 1
    return a


class PolynomialOne:
    """Min-snap scalar polynomial, d2d/trajectory.py:47-82."""
    def __init__(self, Y0, Y1, duration):
        self.duration = duration
        nd = len(Y0)
        no = 2 * nd
        self._der, self._order = nd, no
        self.coefs = np.zeros((nd, no))
        M1 = np.zeros((nd, nd))
        for i in range(nd):
            M1[i, i] = _arr(i, i)
        self.coefs[0, 0:nd] = np.dot(np.linalg.inv(M1), Y0)
        M3 = np.zeros((nd, nd))
        for i in range(nd):
            for j in range(i, nd):
                M3[i, j] = _arr(i, j) * duration ** (j - i)
        M4 = np.zeros((nd, nd))
        for i in range(nd):
            for j in range(nd):
                M4[i, j] = _arr(i, j + nd) * duration ** (j - i + nd)
        M3a0k = np.dot(M3, self.coefs[0, 0:nd])
        self.coefs[0, nd:no] = np.dot(np.linalg.inv(M4), Y1 - M3a0k)
        for d in range(1, nd):
            for pw in range(0, 2 * nd - d):
                self.coefs[d, pw] = _arr(d, pw + d) * self.coefs[0, pw + d]

    def get(self, t):
        Y = np.zeros(self._der)
        for d in range(self._der):
            v = self.coefs[d, -1]
            for j in range(self._order - 2, -1, -1):   # Horner over all 2*nder coefficients, :77-81
                v *= t
                v += self.coefs[d, j]
            Y[d] = v
        return Y


class MinSnap:
    """MinSnapPoly, d2d/trajectory.py:166-187."""
    def __init__(self, Y00=(0, 0), Y10=(1, 0), duration=1.):
        self.duration = duration
        Y0 = np.zeros((2, 4))
        if np.asarray(Y00).ndim == 1: Y0[:, 0] = Y00
        else: Y0 = np.asarray(Y00, float)
        Y1 = np.zeros((2, 4))
        if np.asarray(Y10).ndim == 1: Y1[:, 0] = Y10
        else: Y1 = np.asarray(Y10, float)
        self._polys = [PolynomialOne(Y0[i], Y1[i], duration) for i in range(2)]
        self.t0 = 0

    def reset(self, t0): self.t0 = t0

    def get(self, t):
        return np.array([p.get(t - self.t0) for p in self._polys]).T


class Composite:
    """CompositeTraj, d2d/trajectory.py:190-208: steps[1:] are reset to the previous step's end,
    time wraps with math.fmod, the active step is argmax(steps_end > lapse)."""
    def __init__(self, steps):
        self.steps = steps
        self.steps_dur = [s.duration for s in steps]
        self.steps_end = np.cumsum(self.steps_dur)
        self.duration = np.sum(self.steps_dur)
        for s, st in zip(self.steps[1:], self.steps_end):
            s.reset(st)
        self.t0 = 0.

    def reset(self, t0): self.t0 = t0

    def get(self, t):
        lapse = math.fmod(t - self.t0, self.duration)
        k = int(np.argmax(self.steps_end > lapse))
        return self.steps[k].get(lapse)


class SpaceIndexed:
    """SpaceIndexedTraj, d2d/trajectory.py:220-241: geometry g(lambda) driven by lambda(t)."""
    def __init__(self, geometry, dynamic):
        self.duration = dynamic.duration
        self._geom, self._dyn = geometry, dynamic

    def get(self, t):
        Y = np.zeros((4, 2))
        lam = self._dyn.get(t)
        lam[0] = np.clip(lam[0], 0., 1.)
        g = self._geom.get(lam[0])
        Y[0] = g[0]
        Y[1] = lam[1] * g[1]
        Y[2] = lam[2] * g[1] + lam[1] ** 2 * g[2]
        Y[3] = lam[3] * g[1] + 3 * lam[1] * lam[2] * g[2] + lam[1] ** 3 * g[3]
        return Y


class Tabulated:
    """TrajTabulated, d2d/trajectory_factory.py:149-171: zero-order lookup of a planner solution."""
    def __init__(self, sol_time, sol_x, sol_y, sol_psi, sol_v, wind):
        self.sol_time, self.sol_x, self.sol_y, self.sol_psi, self.sol_v, self.wind = sol_time, sol_x, sol_y, sol_psi, sol_v, wind
        self.t0 = 0.
        self.duration = sol_time[-1]

    def get(self, t):
        Y = np.zeros((4, 2))
        idx = np.argmin(t > self.sol_time)
        Y[0] = self.sol_x[idx], self.sol_y[idx]
        (wx, wy), v, psi = self.wind[idx], self.sol_v[idx], self.sol_psi[idx]
        Y[1] = v * np.cos(psi) + wx, v * np.sin(psi) + wy
        return Y


# named trajectories, d2d/trajectory_factory.py
def traj_two_lines():                                    # :29-37
    s1 = Line([0, 0], [50, 50], v=10., t0=0.)
    return Composite([s1, Line([50, 50], [100, 0], v=10., t0=s1.duration)])


def traj_square():                                       # :40-50
    P = [[0, 0], [50, 0], [50, 50], [0, 50]]
    return Composite([Line(P[k], P[(k + 1) % 4], v=10.) for k in range(4)])


def traj_line_with_intro(Y0=(0, 0), Y1=(0, 50), Y2=(100, 50), r=-25.):     # :53-61
    Yc = (np.asarray(Y0, float) + Y1) / 2
    s1 = Circle(c=Yc, r=r, v=10., alpha0=np.pi / 2, dalpha=np.pi)
    return Composite([s1, Line(Y1, Y2, v=10., t0=s1.duration)])


def traj_with_intro(Y0, traj, duration=8.):              # :65-87
    Y1 = traj.get(0)[0]
    d = np.linalg.norm(Y1 - np.asarray(Y0, float))
    return Composite([Line(Y0, Y1, v=d / duration), traj])


def traj_minsnap_demo():                                 # :110-117
    return MinSnap([[0, 10, 0, 0], [0, 0, 0, 0]], [[200, 0, 0, 0], [200, 10, 0, 0]], duration=33.65)


def traj_si_demo(duration=10.):                          # :177-185
    return SpaceIndexed(Line([0, 20], [100, 20], v=100), PolynomialOne([0, 0.05, 0, 0], [1, 0.05, 0, 0], duration))


# ---------------------------------------------------------------------------
# dynamics, flatness, LQR controller  (d2d/dynamic.py, d2d/guidance.py)
# ---------------------------------------------------------------------------
def cont_dyn(X, U, W, tau_phi=TAU_PHI, tau_v=TAU_V, g=G):
    """Aircraft.cont_dyn, d2d/dynamic.py:14-23 (constant wind, guidance.py:15-16)."""
    x, y, psi, phi, v = X
    return np.array([v * np.cos(psi) + W[0],
                     v * np.sin(psi) + W[1],
                     g / v * np.tan(phi),
                     -1 / tau_phi * (phi - U[0]),
                     -1 / tau_v * (v - U[1])])


def rk4_step(X, U, W, dt, nsub=1, tau_phi=TAU_PHI, tau_v=TAU_V):
    """Fixed-step stand-in for Aircraft.disc_dyn (d2d/dynamic.py:25-28): the reference calls adaptive
    LSODA; north_star defines parity under "the same fixed-step integrator and dt".  Zero-order hold on U,
    `nsub` classical RK4 sub-steps, psi wrapped ONCE at the end of the control step (:27)."""
    X = np.array(X, dtype=float)
    h = dt / nsub
    for _ in range(nsub):
        k1 = cont_dyn(X, U, W, tau_phi, tau_v)
        k2 = cont_dyn(X + 0.5 * h * k1, U, W, tau_phi, tau_v)
        k3 = cont_dyn(X + 0.5 * h * k2, U, W, tau_phi, tau_v)
        k4 = cont_dyn(X + h * k3, U, W, tau_phi, tau_v)
        X = X + (h / 6.0) * (k1 + 2.0 * k2 + 2.0 * k3 + k4)
    X[2] = norm_mpi_pi(X[2])
    return X


def flatness(Ys, W, tau_phi=TAU_PHI, tau_v=TAU_V):
    """DiffFlatness.state_and_input_from_output, d2d/guidance.py:23-47 (phi_dot is never filled, :42-43)."""
    X, U, Xdot = np.zeros(5), np.zeros(2), np.zeros(5)
    X[0], X[1] = Ys[0, 0], Ys[0, 1]
    vax, vay = Ys[1, 0] - W[0], Ys[1, 1] - W[1]
    va2 = vax ** 2 + vay ** 2
    va = np.sqrt(va2)
    X[4] = va
    X[2] = np.arctan2(vay, vax)
    vaxd, vayd = Ys[2, 0], Ys[2, 1]
    Xdot[4] = (vax * vaxd + vay * vayd) / va
    Xdot[2] = (vayd * vax - vaxd * vay) / va2
    X[3] = np.arctan((vayd * vax - vaxd * vay) / va / 9.81)
    U[0] = tau_phi * Xdot[3] + X[3]
    U[1] = tau_v * Xdot[4] + X[4]
    return X, U, Xdot


def cont_jac_3(Xr, g=G):
    """The A[:3,:3], A[:3,3:] blocks of Aircraft.cont_jac (d2d/dynamic.py:32-43) as used at guidance.py:78.
    Two entries are replicated AS WRITTEN: g/va/(1+cos^2 phi) and +g tan(phi)/va^2."""
    psi, phi, va = Xr[2], Xr[3], Xr[4]
    spsi, cpsi = np.sin(psi), np.cos(psi)
    cphi2, tan_phi = np.cos(phi) ** 2, np.tan(phi)
    A1 = np.array([[0., 0., -va * spsi], [0., 0., va * cpsi], [0., 0., 0.]])
    B1 = np.array([[0., cpsi], [0., spsi], [g / va / (1 + cphi2), g / va ** 2 * tan_phi]])
    return A1, B1


def lqr(A, B, Q, R):
    """control.lqr as python-control computes it without slycot (guidance.py:80)."""
    X = scipy.linalg.solve_continuous_are(A, B, Q, R)
    return np.linalg.solve(R, B.T @ X)


_Q, _R = np.diag([1, 1, 0.1]), np.diag([8., 1.])                     # guidance.py:79
_ERR_SATS = np.array([20, 20, np.pi / 3, np.pi / 4, 1])              # guidance.py:69
_U_LO, _U_HI = np.array([-np.deg2rad(45), 4.]), np.array([np.deg2rad(45), 20.])    # guidance.py:87-88


def dfff_control(traj, X, t, W, tau_phi=TAU_PHI, tau_v=TAU_V):
    """DFFFController.get, d2d/guidance.py:62-91.  Returns U, Xr, K(2x3)."""
    Yref = traj.get(t)
    Xr, Ur, _ = flatness(Yref, W, tau_phi, tau_v)
    dX = X - Xr
    dX[2] = norm_mpi_pi(dX[2])
    dX = np.clip(dX, -_ERR_SATS, _ERR_SATS)
    A1, B1 = cont_jac_3(Xr)
    K1 = lqr(A1, B1, _Q, _R)
    U = Ur - K1 @ dX[:3]
    return np.clip(U, _U_LO, _U_HI), Xr, K1


def run_simulation(time, traj, W, X0, perts=None, nsub=1, tau_phi=TAU_PHI, tau_v=TAU_V):
    """run_simulation, 05_test_simulation.py:21-34, with rk4_step in place of LSODA."""
    T = len(time)
    X, U = np.zeros((T, 5)), np.zeros((T, 2))
    Xref, K = np.zeros((T, 5)), np.zeros((T, 2, 3))
    Yref = np.array([traj.get(t) for t in time])
    X[0] = X0
    for i in range(1, T):
        U[i - 1], Xref[i - 1], K[i - 1] = dfff_control(traj, X[i - 1], time[i - 1], W, tau_phi, tau_v)
        X[i] = rk4_step(X[i - 1], U[i - 1], W, time[i] - time[i - 1], nsub, tau_phi, tau_v)
        if perts is not None:
            X[i] += perts[i]
    U[-1], Xref[-1], K[-1] = dfff_control(traj, X[-1], time[-1], W, tau_phi, tau_v)
    return X, U, Yref, Xref, K


# ---------------------------------------------------------------------------
# scenario registry  (d2d/scenario.py) -- parameters only
# ---------------------------------------------------------------------------
def scenario(name):
    """Returns dict(trajs, time, wind, X0s, perts) for the runnable rows of d2d/scenario.py."""
    dt = 0.01                                                           # scenario.py:21
    s = {"perts": None, "X0s": None, "time": None}
    if name == "line":                                                  # :72-85
        s.update(trajs=[Line([0, 25], [100, 25], v=10., t0=0.)], wind=[0., 0.], time=np.arange(0, 12., dt),
                 X0s=[[10, 10, 0, 0, 10]])
        p = np.zeros((len(s["time"]), 5)); p[600, 1] = 10
        s["perts"] = [p]
    elif name == "line2":                                               # :88-98
        s.update(trajs=[traj_two_lines()], wind=[0., 0.], time=np.arange(0, 12., dt), X0s=[[0, 10, 0, 0, 10]])
    elif name == "circle":                                              # :101-118 with cst_gvel=True
        tr = Circle(alpha0=3 * np.pi / 2)
        s.update(trajs=[tr], wind=[5., 0.], time=np.arange(0, tr.duration, dt))
    elif name == "square":                                              # :142-151
        s.update(trajs=[traj_square()], wind=[0., 0.], time=np.arange(0, 30., dt), X0s=[[0, 0, 0, 0, 10]])
    elif name == "mucir":                                               # :155-168
        s.update(trajs=[Circle(c=[40., 50.], alpha0=i * np.deg2rad(30.), v=10.) for i in range(5)],
                 X0s=[[75 - 5 * i, 60 + 5 * i, np.pi, 0, 10] for i in range(5)], wind=[5., 0.],
                 time=np.arange(0, 20, dt))
    elif name == "mucir2":                                              # :172-184
        s.update(trajs=[Circle(c=[40., 50.], alpha0=0.), Circle(c=[40., 50.], alpha0=np.deg2rad(30.))],
                 X0s=[[75, 50, np.pi / 2, 0, 10], [85, 70, np.pi / 1.5, 0, 10]], wind=[1., 0.],
                 time=np.arange(0, 18, dt))
    elif name == "patrol":                                              # :189-205
        s.update(trajs=[traj_line_with_intro([0., 100.], [0., 50.], [200., 50.], 25.),
                        traj_line_with_intro([0., 0.], [0., 50.], [200., 50.], -25.)],
                 X0s=[[0, 100, -np.pi, 0, 10], [0, 0, -np.pi, 0, 10]], wind=[0., 2.5], time=np.arange(0, 20, dt))
    elif name == "patrol_2":                                            # :209-227
        t1 = traj_line_with_intro([0., 100.], [0., 50.], [100., 50.], 25.)
        t2 = traj_with_intro([-20, 0], Slalom(p1=[0, 50], p2=[100, 50], v=10.), duration=8.)
        t3 = traj_with_intro([0, 0], Slalom(p1=[0, 40], p2=[100, 40], v=10.), duration=8.)
        s.update(trajs=[t1, t2, t3], X0s=[[0, 100, -np.pi, 0, 10], [-15, 0, np.pi / 2, 0, 10], [5, 0, np.pi / 2, 0, 10]],
                 wind=[0., 2.5], time=np.arange(0., 17.5, dt))
    elif name == "patrol_3":                                            # :230-248
        trajs = []
        for i in range(2):
            dy = 5 * i; dx = dy / 2
            l1 = Line([0, 10 + dy], [100 - dx, 10 + dy], v=10., t0=0.)
            c1 = Circle(c=[100 - dx, 40], r=30. - dy, v=10., t0=0., alpha0=-np.pi / 2, dalpha=np.pi)
            s2 = Slalom(p1=[100, 60 - dy], p2=[0, 60 - dy], v=10., t0=0., phi=np.pi / 2)
            trajs.append(Composite([l1, c1, s2]))
        s.update(trajs=trajs, wind=[0., 5.])
    elif name == "circForm":                                            # :254-264
        s.update(trajs=[Circle(alpha0=3 * np.pi / 2 + i * np.pi / 6) for i in range(2)], wind=[0., 0.])
    else:
        raise KeyError(name)
    if s["time"] is None:                                               # :29-32
        s["time"] = np.arange(0., np.max([tr.duration for tr in s["trajs"]]), dt)
    if s["perts"] is None:                                              # :35-36
        s["perts"] = [np.zeros((len(s["time"]), 5)) for _ in s["trajs"]]
    if s["X0s"] is None:                                                # :39-45
        s["X0s"] = [flatness(tr.get(s["time"][0]), s["wind"])[0] for tr in s["trajs"]]
    if name == "circForm":                                              # :261-264 (wind set AFTER X0s were built)
        for P0, X0 in zip([[30, 10], [40, 10]], s["X0s"]):
            X0[:2] = P0
        s["wind"] = [0., 5.]
    s["X0s"] = [np.array(x, dtype=float) for x in s["X0s"]]
    return s


# ---------------------------------------------------------------------------
# circular formation  (d2d/guidance.py:99-181, 08_CircularFormation_Full.py:21-97)
# ---------------------------------------------------------------------------
def chain_incidence(n_ac):
    """B of 08_CircularFormation_Full.py:49-60."""
    B = np.zeros((n_ac, n_ac - 1))
    for j in range(n_ac - 1):
        B[j, j], B[j + 1, j] = -1, 1
    return B


def dcf(B, c, p, z_des, kr):
    """DCFController.get, d2d/guidance.py:103-126.  c is (n_ac,2), p is (2,n_ac)."""
    pc = p - c.T
    theta = np.arctan2(pc[1, :], pc[0, :])
    e = B.T @ theta - z_des
    for i in range(len(e)):
        if e[i] > np.pi: e[i] -= 2 * np.pi
        if e[i] <= -np.pi: e[i] += 2 * np.pi
    return -kr * (B @ e), np.rad2deg(e)


def gvf(X, c, r, ke, kd):
    """CircleTraj.get + GVFcontroller.get, d2d/guidance.py:137-146, 155-181."""
    px, py, psi, v = X[0], X[1], X[2], X[4]
    e = ((px - c[0]) ** 2 + (py - c[1]) ** 2) - r ** 2
    n = np.array([2 * (px - c[0]), 2 * (py - c[1])])
    H = np.array([[2., 0.], [0., 2.]])
    E = np.array([[0., 1.], [-1., 0.]])
    pdn = np.array([np.cos(psi), np.sin(psi)])
    p_dot = v * pdn
    tau = E @ n
    pd_dot = tau - ke * e * n
    nrm = np.linalg.norm(pd_dot)
    pd_dot_n = pd_dot / nrm
    o = E @ pd_dot_n
    m = np.outer(o, o)
    mbis = np.outer(n, p_dot)
    U1 = -(m @ ((E - ke * np.identity(2) * e) @ H @ p_dot - ke * mbis @ n))
    U1 = U1 @ (E @ pd_dot_n / nrm)
    U2 = kd * pdn @ E @ pd_dot_n
    return U1 + U2, U1, U2


def run_formation(c, r, n_ac, t_end, ke, kd, kr, z_des, dt=0.05, nsub=5, X1=None, v_c=15., tau_phi=TAU_PHI,
                  tau_v=TAU_V, B=None):
    """Loop of 08_CircularFormation_Full.py:73-94 with c as (n_ac,2) (09_CircularFormation_diffcentre.py:33)."""
    time = np.arange(0, t_end, dt)
    T = len(time)
    X1 = np.array([20, 30, -np.pi / 2, 0, 10.]) if X1 is None else np.asarray(X1, float)
    B = chain_incidence(n_ac) if B is None else np.asarray(B, float)
    X = np.zeros((T, n_ac, 5)); U = np.zeros((T, n_ac)); Rr_log = np.zeros((T, n_ac)); eth = np.zeros((T, B.shape[1]))
    X[0] = X1
    c = np.asarray(c, float)
    for i in range(1, T):
        Ur, e_deg = dcf(B, c, X[i - 1, :, :2].T, z_des, kr)
        Rr = Ur + r
        Rr_log[i], eth[i] = Rr, e_deg
        for j in range(n_ac):
            Ug, _, _ = gvf(X[i - 1, j], c[j], Rr[j], ke, kd)
            U[i - 1, j] = np.arctan(Ug / 9.81)
            X[i, j] = rk4_step(X[i - 1, j], [U[i - 1, j], v_c], [0., 0.], dt, nsub, tau_phi, tau_v)
    return X, U, time, Rr_log, eth


# ---------------------------------------------------------------------------
# 5-state LQR tracker on sampled references  (Controllers.py:50-186, 10_opt_traj_tracking.py:18-90)
# ---------------------------------------------------------------------------
def compute_derivatives(x_ref, y_ref, dt):
    """ComputeDerivatives, 10_opt_traj_tracking.py:18-25."""
    Fdx = np.gradient(x_ref, edge_order=2) / dt
    Fddx = np.gradient(Fdx, edge_order=2) / dt
    Fdy = np.gradient(y_ref, edge_order=2) / dt
    Fddy = np.gradient(Fdy, edge_order=2) / dt
    return Fdx, Fdy, Fddx, Fddy


def flatness5(Y, Yd, Ydd, Yddd, w, tau_phi=TAU_PHI, tau_v=TAU_V, g=G):
    """DiffFlatness.ComputeFlatness, Controllers.py:62-108, formulas as written (note c1's first product
    v_ax*v_ax_dot and the cancelling c2 terms, :95-96)."""
    X, U = np.zeros(5), np.zeros(2)
    v_ax, v_ay = Yd[0] - w[0], Yd[1] - w[1]
    v2 = v_ax ** 2 + v_ay ** 2
    v_ax_dot, v_ay_dot = Ydd[0], Ydd[1]
    v_ax_ddot, v_ay_ddot = Yddd[0], Yddd[1]
    X[0], X[1] = Y[0], Y[1]
    X[2] = np.arctan2(v_ay, v_ax)
    X[3] = np.arctan2((v_ax * v_ay_dot - v_ay * v_ax_dot), g * np.sqrt(v2))
    X[4] = np.sqrt(v2)
    va_dot = (v_ax * v_ax_dot + v_ay * v_ay_dot) / np.sqrt(v2)
    c1 = 1 + ((v_ax * v_ax_dot - v_ax_dot * v_ay) ** 2) / v2
    c2 = v_ax * v_ay_ddot + v_ax_dot * v_ay_dot - v_ax_ddot * v_ay - v_ax_dot * v_ay_dot
    c3 = (v_ax * v_ax_dot + v_ay * v_ay_dot) * (v_ax * v_ay_dot - v_ax_dot * v_ay)
    phi_dot = (1 / c1) * (1 / v2) * (c2 * np.sqrt(v2) - c3 / np.sqrt(v2))
    U[0] = tau_phi * phi_dot + X[3]
    U[1] = tau_v * va_dot + X[4]
    return X, U


def cont_jac_5(Xr, tau_phi=TAU_PHI, tau_v=TAU_V, g=G):
    """Aircraft.cont_jac in full, d2d/dynamic.py:32-43."""
    psi, phi, va = Xr[2], Xr[3], Xr[4]
    spsi, cpsi = np.sin(psi), np.cos(psi)
    cphi2, tan_phi = np.cos(phi) ** 2, np.tan(phi)
    A = np.array([[0., 0., -va * spsi, 0., cpsi],
                  [0., 0., va * cpsi, 0., spsi],
                  [0., 0., 0., g / va / (1 + cphi2), g / va ** 2 * tan_phi],
                  [0., 0., 0., -1 / tau_phi, 0],
                  [0., 0., 0., 0., -1 / tau_v]])
    B = np.array([[0, 0], [0, 0], [0, 0], [1 / tau_phi, 0], [0, 1 / tau_v]])
    return A, B


_Q5, _PHI_LIM5 = np.diag([1, 1, 0.1, 0.01, 0.01]), np.deg2rad(60)        # Controllers.py:149,152


def tracker_gain(X, Y, Yd, Ydd, Yddd, w, tau_phi=TAU_PHI, tau_v=TAU_V):
    """DiffController.ComputeGain, Controllers.py:159-186 -> Xr, dX, U, K (2x5)."""
    Xr, Ur = flatness5(Y, Yd, Ydd, Yddd, w, tau_phi, tau_v)
    dX = X - Xr
    dX[2] = (dX[2] + np.pi) % (2 * np.pi) - np.pi
    dX[3] = (dX[3] + np.pi) % (2 * np.pi) - np.pi
    dX = np.clip(dX, -_ERR_SATS, _ERR_SATS)
    A, B = cont_jac_5(Xr, tau_phi, tau_v)
    K = lqr(A, B, _Q5, _R)
    U = np.clip(Ur - K @ dX, [-_PHI_LIM5, 4], [_PHI_LIM5, 20])
    return Xr, dX, U, K


def run_tracker(time_opt, x_ref, y_ref, w, X0s, nsub=10, tau_phi=TAU_PHI, tau_v=TAU_V):
    """implement_controller, 10_opt_traj_tracking.py:27-90, with rk4_step for LSODA.  x_ref, y_ref: (T, n_ac)."""
    T, n_ac = x_ref.shape
    dt = time_opt[1] - time_opt[0]
    X = np.zeros((T, n_ac, 5)); U = np.zeros((T, n_ac, 2)); Xr_log = np.zeros((T, n_ac, 5)); dX_log = np.zeros((T, n_ac, 5))
    K_log = np.zeros((T - 1, n_ac, 2, 5))
    D = [compute_derivatives(x_ref[:, j], y_ref[:, j], dt) for j in range(n_ac)]
    X[0] = np.asarray(X0s, float)
    for i in range(1, T):
        for j in range(n_ac):
            Fdx, Fdy, Fddx, Fddy = D[j]
            Xr, dX, Uc, K = tracker_gain(X[i - 1, j].copy(), [x_ref[i, j], y_ref[i, j]], [Fdx[i], Fdy[i]], [Fddx[i], Fddy[i]], [0, 0], w, tau_phi, tau_v)
            X[i, j] = rk4_step(X[i - 1, j], Uc, w, dt, nsub, tau_phi, tau_v)
            U[i - 1, j], dX_log[i - 1, j], Xr_log[i - 1, j], K_log[i - 1, j] = Uc, dX, Xr, K
    return X, U, Xr_log, dX_log, K_log


# ---------------------------------------------------------------------------
# collocation  (d2d/opty_utils.py:38-50 EoM, backward Euler as opty's default)
# ---------------------------------------------------------------------------
def planner_timing(t0, t1, hz):
    """d2d/opty_utils.py:8-14."""
    num_nodes = int((t1 - t0) * hz) + 1
    h = 1. / hz
    return num_nodes, h, (num_nodes - 1) * h


def triangle(p0, p1, va, duration, num_nodes, go_left=1.):
    """d2d/opty_utils.py:171-187."""
    p0, p1 = np.asarray(p0, float), np.asarray(p1, float)
    p0p1 = p1 - p0
    d = np.linalg.norm(p0p1)
    u = p0p1 / d; v = np.array([-u[1], u[0]])
    D = va * duration
    p2 = p0 + p0p1 / 2
    if D > d:
        p2 = p2 + np.sign(go_left) * np.sqrt(D ** 2 - d ** 2) / 2 * v
    n1 = int(num_nodes / 2); n2 = num_nodes - n1
    pts = np.vstack((np.linspace(p0, p2, n1), np.linspace(p2, p1, n2)))
    psi0 = np.arctan2((p2 - p0)[1], (p2 - p0)[0]); psi1 = np.arctan2((p1 - p2)[1], (p1 - p2)[0])
    psis = np.hstack((psi0 * np.ones(n1), psi1 * np.ones(n2)))
    return pts[:, 0], pts[:, 1], psis, np.zeros(num_nodes), va * np.ones(num_nodes)


def multi_slices(N, n_ac):
    """Free-vector layout of 07_multioptyplan.py:41-47 (n_ac = 1 gives 06_optyplan.py:35-39)."""
    sx = [slice((0 + 3 * i) * N, (1 + 3 * i) * N) for i in range(n_ac)]
    sy = [slice((1 + 3 * i) * N, (2 + 3 * i) * N) for i in range(n_ac)]
    sp = [slice((2 + 3 * i) * N, (3 + 3 * i) * N) for i in range(n_ac)]
    o = 3 * n_ac * N
    sphi = [slice(o + i * N, o + (i + 1) * N) for i in range(n_ac)]
    o += n_ac * N
    sv = [slice(o + i * N, o + (i + 1) * N) for i in range(n_ac)]
    return sx, sy, sp, sphi, sv


def colloc_residual(free, N, n_ac, h, wind, inst, g=G):
    """Backward-Euler defects of the EoM at d2d/opty_utils.py:42-44 (note `+ w`), ordered equation-major /
    node-minor, followed by the instance constraints `free[k*N+node] - value` (opty layout, appendix B2)."""
    sx, sy, sp, sphi, sv = multi_slices(N, n_ac)
    rows = []
    for a in range(n_ac):
        x, y, psi, phi, v = free[sx[a]], free[sy[a]], free[sp[a]], free[sphi[a]], free[sv[a]]
        rows.append((x[1:] - x[:-1]) / h - v[1:] * np.cos(psi[1:]) + wind[0])
        rows.append((y[1:] - y[:-1]) / h - v[1:] * np.sin(psi[1:]) + wind[1])
        rows.append((psi[1:] - psi[:-1]) / h - g * np.tan(phi[1:]) / v[1:])
    res = np.concatenate(rows)
    iv = np.array([free[int(k) * N + int(node)] - val for (k, node, val) in inst])
    return np.concatenate([res, iv])


def colloc_jac_compact(free, N, n_ac, h, g=G):
    """The 12 structurally non-zero partials per aircraft-node, shape (n_ac, 12, N-1), in
    the order  eq1:[dx_i, dpsi_i, dx_p, dv_i]  eq2:[dy_i, dpsi_i, dy_p, dv_i]  eq3:[dpsi_i, dpsi_p, dphi_i, dv_i]."""
    sx, sy, sp, sphi, sv = multi_slices(N, n_ac)
    out = np.zeros((n_ac, 12, N - 1))
    ih = 1.0 / h
    for a in range(n_ac):
        psi, phi, v = free[sp[a]][1:], free[sphi[a]][1:], free[sv[a]][1:]
        s, c, tn = np.sin(psi), np.cos(psi), np.tan(phi)
        out[a, 0], out[a, 1], out[a, 2], out[a, 3] = ih, v * s, -ih, -c
        out[a, 4], out[a, 5], out[a, 6], out[a, 7] = ih, -v * c, -ih, -s
        out[a, 8], out[a, 9], out[a, 10], out[a, 11] = ih, -ih, -g * (tn ** 2 + 1) / v, g * tn / v ** 2
    return out


# (equation, dense-column) of the 12 compact entries of ONE aircraft inside its own 3 x 8 block
# [x_i, y_i, psi_i, x_p, y_p, psi_p, phi_i, v_i]
_COMPACT_EQ = [0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2]
_COMPACT_LOCAL = [0, 2, 3, 7, 1, 2, 4, 7, 2, 5, 6, 7]


def colloc_structure(N, n_ac, inst, layout="dense"):
    """(rows, cols) of the Jacobian values in opty's COO convention (appendix B3/B4): per node a dense
    n x (2n+q) block, node-major then equation then wrt-variable [x_i.., x_p.., u_i..]; then one entry per
    instance constraint.  layout="compact" lists only the 12 non-zeros per aircraft-node in the order of
    `colloc_jac_compact` flattened as (n_ac, 12, N-1)."""
    n, q = 3 * n_ac, 2 * n_ac
    if layout == "dense":
        i = np.arange(N - 1)[:, None, None]
        e = np.arange(n)[None, :, None]
        cur = [j * N for j in range(n)]; prev = [j * N - 1 for j in range(n)]; inp = [n * N + j * N for j in range(q)]
        base = np.array(cur + prev + inp)[None, None, :]
        rows = np.broadcast_to(e * (N - 1) + i, (N - 1, n, 2 * n + q)).reshape(-1)
        cols = np.broadcast_to(base + i + 1, (N - 1, n, 2 * n + q)).reshape(-1)
    else:
        rows = np.zeros((n_ac, 12, N - 1), dtype=np.int64); cols = np.zeros_like(rows)
        i = np.arange(N - 1)
        for a in range(n_ac):
            colmap = {0: (3 * a) * N + i + 1, 1: (3 * a + 1) * N + i + 1, 2: (3 * a + 2) * N + i + 1,
                      3: (3 * a) * N + i, 4: (3 * a + 1) * N + i, 5: (3 * a + 2) * N + i,
                      6: (n + a) * N + i + 1, 7: (n + n_ac + a) * N + i + 1}
            for k in range(12):
                rows[a, k] = (3 * a + _COMPACT_EQ[k]) * (N - 1) + i
                cols[a, k] = colmap[_COMPACT_LOCAL[k]]
        rows, cols = rows.reshape(-1), cols.reshape(-1)
    ir = n * (N - 1) + np.arange(len(inst)); ic = np.array([int(k) * N + int(node) for (k, node, _) in inst], dtype=np.int64)
    return np.concatenate([rows, ir]).astype(np.int64), np.concatenate([cols, ic]).astype(np.int64)


def colloc_jac_dense(free, N, n_ac, h, n_inst=0):
    """opty-dense Jacobian values: (N-1, 3 n_ac, 8 n_ac) flattened, then `n_inst` ones."""
    n, q = 3 * n_ac, 2 * n_ac
    comp = colloc_jac_compact(free, N, n_ac, h)
    J = np.zeros((N - 1, n, 2 * n + q))
    for a in range(n_ac):
        dense_col = [3 * a, 3 * a + 1, 3 * a + 2, n + 3 * a, n + 3 * a + 1, n + 3 * a + 2, 2 * n + a, 2 * n + n_ac + a]
        for k in range(12):
            J[:, 3 * a + _COMPACT_EQ[k], dense_col[_COMPACT_LOCAL[k]]] = comp[a, k]
    return np.concatenate([J.reshape(-1), np.ones(n_inst)])


# cost specification: a dict with the weights of the reference's cost classes
def cost_and_grad(free, N, n_ac, spec, multi=None):
    """Cost classes of d2d/opty_utils.py:55-165 (multi=False, the single-aircraft planner) and
    d2d/multiopty_utils.py:29-174 (multi=True).  `spec` keys (all optional):
      obj_scale, vsp, kvel, kbank          -> CostInput / CostAirVel / CostBank terms
      obstacles=[(cx,cy,r),..], kobs, obs_kind (aircraft 0 only in the multi classes, :74)
      kcol, rcol, kcol_k=2, pairs='01'|'all' -> CostCollision (:120-153; 'all' is the all-pairs generalisation
                                               of SURVEY D11), exact_grad=False replicates the reference's
                                               gradients that omit (k/r)^2.
    Normalisation: single-aircraft classes divide by N; multi CostInput divides by N*n_ac, obstacle and
    collision by N (:38,:62,:90,:134)."""
    multi = (n_ac > 1) if multi is None else multi
    sx, sy, sp, sphi, sv = multi_slices(N, n_ac)
    s = spec.get("obj_scale", 1.)
    vsp, kvel, kbank = spec.get("vsp", 10.), spec.get("kvel", 0.), spec.get("kbank", 0.)
    exact = spec.get("exact_grad", False)
    grad = np.zeros_like(free)
    cost = 0.
    norm_in = s / N / n_ac if multi else s / N
    for a in range(n_ac):
        phi, v = free[sphi[a]], free[sv[a]]
        cost += norm_in * (kvel * np.sum(np.square(v - vsp)) + kbank * np.sum(np.square(phi)))
        grad[sphi[a]] += norm_in * kbank * 2 * phi
        grad[sv[a]] += norm_in * kvel * 2 * (v - vsp)
    kobs = spec.get("kobs", float("nan"))
    if not np.isnan(kobs):
        kind = spec.get("obs_kind", 0)
        x, y = free[sx[0]], free[sy[0]]
        for (cx, cy, r) in spec.get("obstacles", ()):
            dx, dy = x - cx, y - cy
            if kind == 0:
                with np.errstate(over="ignore"):
                    es = np.clip(np.exp(r ** 2 - (np.square(dx) + np.square(dy))), 0., 1e3)
                f = 1.
            else:
                k = 2.
                es = np.exp(-(np.square(dx / r * k) + np.square(dy / r * k)))
                f = (k / r) ** 2 if exact else 1.
            cost += kobs * s / N * np.sum(es)
            grad[sx[0]] += kobs * (s / N * -2. * dx * es) * f
            grad[sy[0]] += kobs * (s / N * -2. * dy * es) * f
    kcol = spec.get("kcol", float("nan"))
    if not np.isnan(kcol):
        r, k = spec.get("rcol", 3.), spec.get("kcol_k", 2.)
        f = (k / r) ** 2 if exact else 1.
        pairs = [(0, 1)] if spec.get("pairs", "01") == "01" else [(a, b) for a in range(n_ac) for b in range(a + 1, n_ac)]
        for (a, b) in pairs:
            dx, dy = free[sx[a]] - free[sx[b]], free[sy[a]] - free[sy[b]]
            es = np.exp(-(np.square(dx / r * k) + np.square(dy / r * k)))
            cost += kcol * s / N * np.sum(es)
            grad[sx[a]] += kcol * (s / N * -2. * dx * es) * f
            grad[sy[a]] += kcol * (s / N * -2. * dy * es) * f
            grad[sx[b]] += kcol * (s / N * 2. * dx * es) * f
            grad[sy[b]] += kcol * (s / N * 2. * dy * es) * f
    return cost, grad


# ---------------------------------------------------------------------------
# single shooting on the collocation grid (checker of d2dx_shoot_forward / d2dx_shoot_adjoint)
# ---------------------------------------------------------------------------
def shoot_states(phi, v, p0, h, wind, g=G):
    """States that zero the backward-Euler defects of `colloc_residual` for given inputs.  phi, v: (n_ac, N)
    (node 0 unused); p0: (3, n_ac).  Returns x, y, psi of shape (n_ac, N)."""
    phi, v, p0 = np.atleast_2d(phi), np.atleast_2d(v), np.asarray(p0, float).reshape(3, -1)
    z = np.zeros((phi.shape[0], 1))
    psi = p0[2][:, None] + np.hstack([z, np.cumsum(h * g * np.tan(phi[:, 1:]) / v[:, 1:], axis=1)])
    x = p0[0][:, None] + np.hstack([z, np.cumsum(h * (v[:, 1:] * np.cos(psi[:, 1:]) - wind[0]), axis=1)])
    y = p0[1][:, None] + np.hstack([z, np.cumsum(h * (v[:, 1:] * np.sin(psi[:, 1:]) - wind[1]), axis=1)])
    return x, y, psi


def shoot_free(phi, v, p0, h, wind):
    """Planner-layout free vector [x, y, psi per aircraft | phi per aircraft | v per aircraft] of a shooting point."""
    x, y, psi = shoot_states(phi, v, p0, h, wind)
    return np.concatenate([np.concatenate([x[a], y[a], psi[a]]) for a in range(x.shape[0])] + [np.ravel(phi), np.ravel(v)])


def shoot_lagrangian(phi, v, p0, p1, h, wind, spec, lam, rho, multi=None, state_box=None):
    """cost (reference cost classes, exact value; plus the soft state box weight * obj_scale / N * sum of squared excesses
    when state_box = (x_lo, x_hi, y_lo, y_hi, weight)), c = terminal state - p1 (3, n_ac), and cost + sum lam c + rho/2 |c|^2."""
    phi, v = np.atleast_2d(phi), np.atleast_2d(v)
    n_ac, N = phi.shape
    x, y, psi = shoot_states(phi, v, p0, h, wind)
    cost, _ = cost_and_grad(shoot_free(phi, v, p0, h, wind), N, n_ac, spec, multi)
    if state_box is not None:
        xl, xh, yl, yh, wgt = state_box
        ex, ey = np.maximum(x - xh, 0) + np.minimum(x - xl, 0), np.maximum(y - yh, 0) + np.minimum(y - yl, 0)
        cost += wgt * spec.get("obj_scale", 1.) / N * np.sum(ex ** 2 + ey ** 2)
    c = np.stack([x[:, -1], y[:, -1], psi[:, -1]]) - np.asarray(p1, float).reshape(3, -1)
    return cost, c, cost + np.sum(np.asarray(lam) * c) + 0.5 * rho * np.sum(c * c)


def shoot_value_and_grad(phi, v, p0, p1, h, wind, spec, lam, rho, multi=None, g=G):
    """Lagrangian of `shoot_lagrangian` and its exact gradient with respect to the inputs by the adjoint recursion,
    vectorised with cumulative sums -- for the input-cost terms only (kvel, kbank; no obstacle / collision / box terms).
    phi, v: (n_ac, N).  Returns L, dL/dphi (n_ac, N), dL/dv (n_ac, N), cost, c."""
    phi, v = np.atleast_2d(np.asarray(phi, float)), np.atleast_2d(np.asarray(v, float))
    n_ac, N = phi.shape
    multi = (n_ac > 1) if multi is None else multi
    x, y, psi = shoot_states(phi, v, p0, h, wind)
    s = spec.get("obj_scale", 1.)
    vsp, kvel, kbank = spec.get("vsp", 10.), spec.get("kvel", 0.), spec.get("kbank", 0.)
    norm_in = s / N / n_ac if multi else s / N
    cost = norm_in * (kvel * np.sum(np.square(v - vsp)) + kbank * np.sum(np.square(phi)))
    c = np.stack([x[:, -1], y[:, -1], psi[:, -1]]) - np.asarray(p1, float).reshape(3, -1)
    lam = np.asarray(lam, float).reshape(3, -1)
    L = cost + np.sum(lam * c) + 0.5 * rho * np.sum(c * c)
    gl = lam + rho * c                                            # dL / d terminal state, (3, n_ac)
    # x_N = x_0 + sum_j h (v_j cos psi_j - w): d x_N / d psi_j = -h v_j sin psi_j, and psi_j depends on (phi_i, v_i) for i <= j
    sx, sy = -h * v[:, 1:] * np.sin(psi[:, 1:]), h * v[:, 1:] * np.cos(psi[:, 1:])
    Gpsi = np.cumsum((gl[0][:, None] * sx + gl[1][:, None] * sy)[:, ::-1], axis=1)[:, ::-1] + gl[2][:, None]
    dphi, dv = np.zeros_like(phi), np.zeros_like(v)
    dphi[:, 1:] = Gpsi * h * g / (np.cos(phi[:, 1:]) ** 2 * v[:, 1:])
    dv[:, 1:] = Gpsi * (-h * g * np.tan(phi[:, 1:]) / v[:, 1:] ** 2) + gl[0][:, None] * h * np.cos(psi[:, 1:]) + gl[1][:, None] * h * np.sin(psi[:, 1:])
    dphi += norm_in * kbank * 2 * phi
    dv += norm_in * kvel * 2 * (v - vsp)
    return L, dphi, dv, cost, c
