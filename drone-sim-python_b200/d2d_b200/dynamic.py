"""Planar aircraft model -- the `d2d.dynamic.Aircraft` call surface (d2d/dynamic.py:5-43), evaluated by the
engine.  Every method accepts one state (shape (5,)) or a batch (shape (n, 5)); `disc_dyn` integrates with
fixed-step RK4 (`nsub` sub-steps, zero-order hold) where the reference calls adaptive LSODA."""
import numpy as np

from .engine import get_engine


def _soa(eng, a, width):
    """(width,) or (n, width) host array -> device [width][n], plus whether the input was a single vector."""
    a = np.asarray(a, dtype=np.float64)
    single = a.ndim == 1
    a2 = a.reshape(1, width) if single else a
    return eng.to_device(np.ascontiguousarray(a2.T)), single


def _wind_soa(eng, W, t, X, n):
    w = np.asarray(W.sample(t, np.asarray(X)[..., :2]) if hasattr(W, "sample") else W, dtype=np.float64)
    return eng.to_device(np.ascontiguousarray(np.broadcast_to(w.reshape(-1, 2) if w.ndim > 1 else w, (n, 2)).T))


class Aircraft:
    i_phi, i_va, i_size = 0, 1, 2
    s_x, s_y, s_psi, s_phi, s_va, s_size = 0, 1, 2, 3, 4, 5
    s_slice_pos = slice(s_x, s_y + 1)
    g = 9.81

    def __init__(self, tau_phi=0.01, tau_v=1., nsub=1):
        self.tau_phi, self.tau_v = tau_phi, tau_v            # d2d/dynamic.py:11-12
        self.nsub = nsub                                     # RK4 sub-steps per control step

    def _ac(self, eng, n):
        return eng.to_device(np.ascontiguousarray(np.broadcast_to(np.array([[self.tau_phi], [self.tau_v]]), (2, n))))

    def cont_dyn(self, X, t, U, W):
        """Xdot = f(X, U) (d2d/dynamic.py:14-23).  Returns a list of 5 floats for one state, (n,5) for a batch."""
        eng = get_engine()
        Xd, single = _soa(eng, X, 5)
        Ud, _ = _soa(eng, U, 2)
        n = Xd.shape[1]
        out = eng.cont_dyn(Xd, Ud, _wind_soa(eng, W, t, X, n), self._ac(eng, n)).cpu().numpy().T
        return list(out[0]) if single else out

    def disc_dyn(self, Xk, Uk, W, t, dt):
        """One zero-order-hold step with psi wrapped to (-pi, pi] (d2d/dynamic.py:25-28)."""
        eng = get_engine()
        Xd, single = _soa(eng, Xk, 5)
        Ud, _ = _soa(eng, np.asarray(Uk, dtype=np.float64).reshape(-1, 2) if not single else np.asarray(Uk, dtype=np.float64).reshape(2), 2)
        n = Xd.shape[1]
        out = eng.disc_dyn(Xd, Ud, _wind_soa(eng, W, t, Xk, n), self._ac(eng, n), dt, self.nsub).cpu().numpy().T
        return out[0] if single else out

    def cont_jac(self, Xr, Ur, t, W):
        """Linearisation (A 5x5, B 5x2) at the reference state, entries as the reference writes them
        (d2d/dynamic.py:32-43)."""
        eng = get_engine()
        Xd, single = _soa(eng, Xr, 5)
        n = Xd.shape[1]
        A, Bm = eng.cont_jac(Xd, self._ac(eng, n))
        A = A.cpu().numpy().T.reshape(n, 5, 5); Bm = Bm.cpu().numpy().T.reshape(n, 5, 2)
        return (A[0], Bm[0]) if single else (A, Bm)
