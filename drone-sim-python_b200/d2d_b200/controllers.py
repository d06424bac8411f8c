"""Full-state LQR tracker on sampled references -- the call surface of the reference's `Controllers.py`
(DiffFlatness.ComputeFlatness :62-108, DiffController.ComputeGain :159-186) and of `implement_controller`
(10_opt_traj_tracking.py:27-90), evaluated by the engine (csrc/d2dx_tracker.cu).

Single calls keep the reference's signatures; `implement_controller` runs the whole tracking loop of every aircraft
in one kernel launch, `track` is its array-level form (M aircraft, T samples)."""
import numpy as np
import torch

from .engine import get_engine


def ComputeDerivatives(x_ref, y_ref, dt):
    """First and second time derivatives of a sampled path by second-order finite differences
    (10_opt_traj_tracking.py:18-25).  Host-side preparation of the reference table."""
    Fdx = np.gradient(x_ref, edge_order=2) / dt
    Fddx = np.gradient(Fdx, edge_order=2) / dt
    Fdy = np.gradient(y_ref, edge_order=2) / dt
    Fddy = np.gradient(Fdy, edge_order=2) / dt
    return Fdx, Fdy, Fddx, Fddy


class CircleTraj:
    """Controllers.py:17-32: circle reference sampled at one time (only used from commented-out code upstream; the tracker
    follows tabulated planner output).  As written: the third derivative carries no omega^3 factor."""

    def __init__(self, v, r=40, c=[0, 0]):
        self.c, self.r, self.v = c, r, v
        self.omega = self.v / self.r

    def TrajPoints(self, t):
        th, r, om = self.omega * t, self.r, self.omega
        Y_ref = [r * np.cos(th), r * np.sin(th)]
        Yd_ref = [-r * np.sin(th) * om, r * np.cos(th) * om]
        Ydd_ref = [-r * np.cos(th) * om ** 2, -r * np.sin(th) * om ** 2]
        Yddd_ref = [r * np.sin(th), -r * np.cos(th)]
        return Y_ref, Yd_ref, Ydd_ref, Yddd_ref


class DiffFlatness:
    """Reference state and input from the flat output and three derivatives (Controllers.py:50-108)."""

    def __init__(self, w=[0, 0], tau_phi=0.01, tau_v=1.):
        self.w, self.g = w, 9.81
        self.tau_phi, self.tau_v = tau_phi, tau_v            # the reference builds a default Aircraft() here (:101)
        self.x_i, self.y_i, self.psi_i, self.phi_i, self.v_i = range(5)

    def ComputeFlatness(self, t, Y_ref, Yd_ref, Ydd_ref, Yddd_ref):
        eng = get_engine()
        Ys = np.concatenate([np.asarray(a, dtype=np.float64).reshape(2) for a in (Y_ref, Yd_ref, Ydd_ref, Yddd_ref)]).reshape(8, 1)
        W = np.asarray(self.w, dtype=np.float64).reshape(2, 1)
        ac = np.array([[self.tau_phi], [self.tau_v]])
        Xr, Ur = eng.flatness5(eng.to_device(Ys), eng.to_device(W), eng.to_device(ac))
        return Xr.cpu().numpy()[:, 0], Ur.cpu().numpy()[:, 0]


class DiffController:
    """Feed-forward + full-state LQR feedback (Controllers.py:139-186)."""

    def __init__(self, w=[0, 0]):
        self.w = w
        self.DF = DiffFlatness(self.w)
        self.psi_i, self.phi_i = 2, 3
        self.err_sats = np.array([20, 20, np.pi / 3, np.pi / 4, 1])
        self.v_min, self.v_max = 4, 20
        self.phi_lim = np.deg2rad(60)
        self.Q, self.R = [1, 1, 0.1, 0.01, 0.01], [8, 1]
        self.K = []
        self._state = None

    def RestrictAngle(self, theta):
        from .guidance import norm_mpi_pi
        return norm_mpi_pi(theta)

    def _gains(self, eng):
        from . import _lib
        g = _lib.TrackerGains()
        g.q[:] = self.Q; g.r[:] = self.R; g.err_sat[:] = list(self.err_sats)
        g.u_lo[:] = [-self.phi_lim, self.v_min]; g.u_hi[:] = [self.phi_lim, self.v_max]
        return g

    def ComputeGain(self, t, X, Y_ref, Yd_ref, Ydd_ref, Yddd_ref, ac):
        eng = get_engine()
        if self._state is None:
            self._state = eng.zeros(7, 1)
        Ys = np.concatenate([np.asarray(a, dtype=np.float64).reshape(2) for a in (Y_ref, Yd_ref, Ydd_ref, Yddd_ref)]).reshape(8, 1)
        U, Xr, dX, K = eng.tracker_control(eng.to_device(np.asarray(X, dtype=np.float64).reshape(5, 1)), eng.to_device(Ys),
                                           eng.to_device(np.asarray(self.w, dtype=np.float64).reshape(2, 1)),
                                           eng.to_device(np.array([[ac.tau_phi], [ac.tau_v]])), gains=self._gains(eng),
                                           lqr_state=self._state)
        self.K.append(K.cpu().numpy()[:, 0].reshape(2, 5))
        return Xr.cpu().numpy()[:, 0], dX.cpu().numpy()[:, 0], U.cpu().numpy()[:, 0]


def track(time_opt, x_ref, y_ref, w, X0s, nsub=10, tau_phi=0.01, tau_v=1., ctrl=None, engine=None, return_gain=False):
    """Tracking loop for M aircraft and T samples: x_ref, y_ref (T, M); w (2,) or (M, 2); X0s (M, 5).
    Returns X (T,M,5), U (T,M,2), X_ref (T,M,5), Yd (T,M,2), Ydd (T,M,2), dX (T,M,5) [, K (T,M,2,5)] with the row
    conventions of implement_controller (row i-1 of U / X_ref / dX belongs to the step i-1 -> i)."""
    eng = engine or get_engine()
    time_opt = np.asarray(time_opt, dtype=np.float64)
    x_ref, y_ref = np.asarray(x_ref, dtype=np.float64), np.asarray(y_ref, dtype=np.float64)
    T, M = x_ref.shape
    dt = time_opt[1] - time_opt[0]                                               # :42
    ref = np.zeros((T, 6, M))
    for j in range(M):
        Fdx, Fdy, Fddx, Fddy = ComputeDerivatives(x_ref[:, j], y_ref[:, j], dt)
        ref[:, 0, j], ref[:, 1, j], ref[:, 2, j], ref[:, 3, j], ref[:, 4, j], ref[:, 5, j] = x_ref[:, j], y_ref[:, j], Fdx, Fdy, Fddx, Fddy
    wind = np.broadcast_to(np.asarray(w, dtype=np.float64).reshape(-1, 2), (M, 2))
    ac = np.stack([np.broadcast_to(np.asarray(tau_phi, dtype=np.float64), (M,)), np.broadcast_to(np.asarray(tau_v, dtype=np.float64), (M,))])
    X_log, U_log = eng.zeros(T, 5, M), eng.zeros(T, 2, M)
    Xr_log, dX_log = eng.zeros(T, 5, M), eng.zeros(T, 5, M)
    K_log = eng.zeros(T, 10, M) if return_gain else None
    flags = eng.zeros(M, dtype=torch.int32)
    gains = (ctrl or DiffController(w))._gains(eng)
    eng.rollout_tracker(eng.to_device(ref), eng.to_device(np.ascontiguousarray(np.asarray(X0s, dtype=np.float64).reshape(M, 5).T)),
                        eng.to_device(np.ascontiguousarray(wind.T)), eng.to_device(ac), dt, 0, T - 1, nsub, gains=gains,
                        X_log=X_log, U_log=U_log, Xr_log=Xr_log, dX_log=dX_log, K_log=K_log, flags=flags)
    tr = lambda a: a.permute(0, 2, 1).contiguous().cpu().numpy()
    out = [tr(X_log), tr(U_log), tr(Xr_log), ref[:, 2:4].transpose(0, 2, 1).copy(), ref[:, 4:6].transpose(0, 2, 1).copy(), tr(dX_log)]
    # Yd / Ydd arrays of the reference are filled for rows 0..T-2 with the samples 1..T-1 (:88-89)
    out[3] = np.concatenate([out[3][1:], np.zeros((1, M, 2))]); out[4] = np.concatenate([out[4][1:], np.zeros((1, M, 2))])
    if return_gain:
        out.append(tr(K_log).reshape(T, M, 2, 5))
    out.append(flags.cpu().numpy())
    return out


def implement_controller(n_ac, df, v, w, X0s, nsub=10):
    """Drop-in for implement_controller(n_ac, df, v, w, X0s) (10_opt_traj_tracking.py:27-90) ->
    X_array, U_array, X_ref_array, Yd_ref_array, Ydd_ref_array, time_opt, dX_array.  `df` is the planner's CSV as a
    DataFrame (columns time, x_i, y_i, psi_i, ...)."""
    time_opt = np.array(df["time"])
    x_ref = np.stack([np.array(df[f"x_{i + 1}"]) for i in range(n_ac)], 1)
    y_ref = np.stack([np.array(df[f"y_{i + 1}"]) for i in range(n_ac)], 1)
    X, U, Xr, Yd, Ydd, dX, _ = track(time_opt, x_ref, y_ref, w, np.asarray(X0s, dtype=np.float64), nsub=nsub)
    return X, U, Xr, Yd, Ydd, time_opt, dX
