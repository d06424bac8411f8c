"""ctypes driver of the host build of the planner's second-order solver (csrc/host_check.cu -> tests/native/libd2dx_hostcheck.so):
used by tests/test_ddp_cpu.py and for development without a GPU."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "tests", "native", "libd2dx_hostcheck.so")


class DdpOptions(C.Structure):
    _fields_ = [("max_iter", C.c_int32), ("max_outer", C.c_int32), ("max_inner", C.c_int32), ("ls_max", C.c_int32), ("ctol", C.c_double),
                ("rel_tol", C.c_double), ("abs_tol", C.c_double), ("rho0", C.c_double), ("rho_growth", C.c_double), ("rho_max", C.c_double),
                ("mu0", C.c_double), ("mu_min", C.c_double), ("mu_max", C.c_double), ("mu_factor", C.c_double), ("reg_mode", C.c_int32), ("min_solved", C.c_int32)]


def default_options(**kw):
    o = DdpOptions(400, 30, 40, 12, 1e-8, 1e-10, 1e-14, 10., 10., 1e8, 1e-6, 1e-8, 1e10, 1.6, 0, 0)
    for k, v in kw.items():
        setattr(o, k, v)
    return o


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB)
        dp = C.POINTER(C.c_double)
        _lib.d2dx_host_ddp_solve.argtypes = [dp, C.c_int, dp, dp, dp, dp, dp, C.POINTER(DdpOptions), dp, dp, dp]
    return _lib


def solve(N, h, wind, vsp, kvel, kbank, bounds, z0, zt, phi0, v0, obstacles=(), kobs=0., obs_kind=0, box=None, obj_scale=1., in_div=1, opts=None):
    """one problem; weights are the planner's (kvel, kbank, kobs), normalised here like d2dx_ddp.cu does.  Returns u (2,N), xs (3,N), info dict"""
    dp = C.POINTER(C.c_double)
    sN = obj_scale / N
    prob = np.array([N, h, wind[0], wind[1], vsp, kvel * sN / in_div, kbank * sN / in_div, kobs * sN if len(obstacles) else 0., obs_kind], float)
    obs = np.ascontiguousarray(np.asarray(obstacles, float).reshape(-1, 3)) if len(obstacles) else np.zeros((1, 3))
    b = np.asarray(bounds, float)
    bx = None if box is None else np.array([box[0], box[1], box[2], box[3], box[4] * sN], float)
    u = np.ascontiguousarray(np.stack([np.broadcast_to(phi0, (N,)), np.broadcast_to(v0, (N,))]).astype(float))
    xs, info = np.zeros((3, N)), np.zeros(8)
    z0, zt = np.asarray(z0, float), np.asarray(zt, float)
    o = opts or default_options()
    P = lambda a: a.ctypes.data_as(dp)
    lib().d2dx_host_ddp_solve(P(prob), len(obstacles), P(obs), P(b), None if bx is None else P(bx), P(z0), P(zt), C.byref(o), P(u), P(xs), P(info))
    keys = ("flag", "iterations", "outer", "cost", "cmax", "lagr", "mu", "rho")
    return u, xs, dict(zip(keys, info))


if __name__ == "__main__":
    import time
    # exp_0: turn around, 10 s at 10 Hz; and the C3 grid (20 s at 50 Hz)
    for N, h, T in ((101, 0.1, 10.), (1001, 0.02, 20.)):
        for phi0 in (0.1, -0.1, 0.3):
            t0 = time.perf_counter()
            u, xs, info = solve(N, h, (0., 0.), 12., 1., 0., (-np.deg2rad(30), np.deg2rad(30), 9., 14.), (0., 0., 0.), (0., 30., np.pi), phi0, 12.)
            print(N, phi0, {k: (f"{v:.3e}" if isinstance(v, float) and abs(v) < 1e-2 else v) for k, v in info.items()}, f"{time.perf_counter() - t0:.3f}s")
