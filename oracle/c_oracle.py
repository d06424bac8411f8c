"""ctypes view of oracle/_build/liboracle.so (the C restatement of the closed-loop oracle).
TEST INFRASTRUCTURE: importable from tests/, __graft_entry__.smoke() and bench.py's CPU legs only."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
T_LINE, T_CIRCLE, T_POLY = 0, 1, 3


def load():
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "d2d_oracle.c")):
        subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)
    lib = C.CDLL(_SO)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    lib.orc_rollout_dfff.restype = C.c_int
    lib.orc_rollout_dfff.argtypes = [C.c_int, C.c_int, dp, ip, dp, dp, dp, C.c_double, C.c_double, C.c_int, C.c_int,
                                     dp, dp, dp, dp, dp, dp, C.c_int]
    lib.orc_lqr3.restype = C.c_int
    lib.orc_lqr3.argtypes = [dp, dp, dp]
    lib.orc_max_threads.restype = C.c_int
    return lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def circle_par(cx, cy, r, v, alpha0, t0=0.):
    B = len(cx)
    par = np.zeros((B, 17))
    par[:, 0], par[:, 1], par[:, 2], par[:, 3], par[:, 4], par[:, 5] = t0, cx, cy, r, np.asarray(v) / np.asarray(r), alpha0
    return np.full(B, T_CIRCLE, np.int32), par


def rollout(time, types, par, wind, X0, tau_phi=0.01, tau_v=1., nsub=1, log_every=1, want_log=True, want_K=False, nthreads=0):
    """Returns dict(X (B,n_rows,5), U (B,n_rows,2), K, X_final, sum_sq_err, max_err, failed)."""
    lib = load()
    time = np.ascontiguousarray(time, np.float64); T = len(time)
    types = np.ascontiguousarray(types, np.int32); par = np.ascontiguousarray(par, np.float64)
    X0 = np.ascontiguousarray(X0, np.float64).reshape(-1, 5); B = len(X0)
    wind = np.ascontiguousarray(np.broadcast_to(np.asarray(wind, np.float64).reshape(-1, 2), (B, 2)))
    n_rows = (T - 1) // log_every + 1
    X = np.zeros((B, n_rows, 5)) if want_log else None
    U = np.zeros((B, n_rows, 2)) if want_log else None
    K = np.zeros((B, n_rows, 6)) if want_K else None
    Xf, ss, mx = np.zeros((B, 5)), np.zeros(B), np.zeros(B)
    bad = lib.orc_rollout_dfff(B, T, _p(time), types.ctypes.data_as(C.POINTER(C.c_int)), _p(par), _p(wind), _p(X0),
                               tau_phi, tau_v, nsub, log_every, _p(X), _p(U), _p(K), _p(Xf), _p(ss), _p(mx), nthreads)
    return {"X": X, "U": U, "K": K, "X_final": Xf, "sum_sq_err": ss, "max_err": mx, "failed": bad}


def lqr3(A, B):
    lib = load()
    A = np.ascontiguousarray(A, np.float64); B = np.ascontiguousarray(B, np.float64); K = np.zeros(6)
    rc = lib.orc_lqr3(_p(A), _p(B), _p(K))
    return K.reshape(2, 3), rc


def max_threads():
    return load().orc_max_threads()
