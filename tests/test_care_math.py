"""The engine's per-step Riccati reductions on the CPU: csrc/host_check.cu compiles care_gain (3-state, d2dx_device.cuh) and
lqr5_gain (5-state, d2dx_lqr5.cuh) -- the very source the kernels inline -- for the host, and this test compares them with
scipy.linalg.solve_continuous_are on the pair the reference builds (Aircraft.cont_jac, d2d/dynamic.py:32-43, both
"as written" entries; control.lqr = CARE + R^-1 B^T P, d2d/guidance.py:78-82, Controllers.py:174)."""
import ctypes as C
import os

import numpy as np
import pytest
import scipy.linalg as sla

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "native", "libd2dx_hostcheck.so")
G = 9.81


@pytest.fixture(scope="module")
def host():
    if not os.path.exists(LIB):
        pytest.fail(f"{LIB} is missing: build it with `python __graft_entry__.py` (make -C drone-sim-python_b200/csrc)")
    lib = C.CDLL(LIB)
    dp = C.POINTER(C.c_double)
    lib.d2dx_host_care_gain.argtypes = [dp, C.c_double, C.c_double, C.c_int, dp, dp]
    lib.d2dx_host_lqr5_gain.argtypes = [dp, dp, C.c_double, C.c_double, C.c_double, C.c_double, dp, dp]
    return lib


def _arr(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


def scipy_gain3(v, phi, q, r):
    """K1 of DFFFController.get at psi_ref = 0."""
    A = np.zeros((3, 3)); A[1, 2] = v
    B = np.array([[0., 1.], [0., 0.], [G / v / (1 + np.cos(phi) ** 2), G * np.tan(phi) / v ** 2]])
    Q, R = np.diag([q[0], q[0], q[1]]), np.diag(r)
    P = sla.solve_continuous_are(A, B, Q, R)
    return np.linalg.solve(R, B.T @ P)


def scipy_gain5(v, phi, q, r, tau_phi, tau_v):
    """K of DiffController.ComputeGain at psi_ref = 0 (full 5-state pair of cont_jac)."""
    A = np.zeros((5, 5)); B = np.zeros((5, 2))
    A[0, 4] = 1.; A[1, 2] = v
    A[2, 3] = G / v / (1 + np.cos(phi) ** 2); A[2, 4] = G * np.tan(phi) / v ** 2
    A[3, 3] = -1. / tau_phi; A[4, 4] = -1. / tau_v
    B[3, 0] = 1. / tau_phi; B[4, 1] = 1. / tau_v
    Q, R = np.diag(q), np.diag(r)
    P = sla.solve_continuous_are(A, B, Q, R)
    return np.linalg.solve(R, B.T @ P)


def test_reduced_3x3_riccati_cold_start_matches_scipy(host):
    rng = np.random.default_rng(0)
    worst = 0.
    for _ in range(400):
        v, phi = rng.uniform(0.3, 60.), rng.uniform(-1.4, 1.4)
        q = (rng.uniform(0.2, 5.), rng.uniform(0.02, 2.)); r = (rng.uniform(0.5, 20.), rng.uniform(0.2, 5.))
        qr, pqr = _arr([q[0], q[1], r[0], r[1]])
        st, pst = _arr(np.zeros(5)); K0, pK = _arr(np.zeros(6))
        assert host.d2dx_host_care_gain(pqr, v, phi, 1, pst, pK) == 1
        Ks = scipy_gain3(v, phi, q, r)
        worst = max(worst, np.abs(K0.reshape(2, 3) - Ks).max() / np.abs(Ks).max())
    assert worst < 1e-10, worst


def test_reduced_3x3_riccati_warm_path_along_a_smooth_reference(host):
    """The rollout's straight-line path (extrapolate, one Newton step, one chord step) on a slowly varying reference, warm
    started from the previous sample, including a jump (trajectory corner) that must fall back to the loop."""
    qr, pqr = _arr([1., 0.1, 8., 1.])
    st, pst = _arr(np.zeros(5)); K0, pK = _arr(np.zeros(6))
    t = np.arange(0, 20, 0.01)
    v = 10 + 3 * np.sin(0.3 * t); phi = 0.4 * np.sin(0.5 * t)
    v[1200:] += 6.; phi[1200:] -= 0.5                          # corner
    worst = 0.
    for k in range(len(t)):
        assert host.d2dx_host_care_gain(pqr, v[k], phi[k], 1 if k == 0 else 0, pst, pK) == 1
        if k % 7 == 0 or 1195 <= k <= 1210:
            Ks = scipy_gain3(v[k], phi[k], (1., 0.1), (8., 1.))
            worst = max(worst, np.abs(K0.reshape(2, 3) - Ks).max() / np.abs(Ks).max())
    assert worst < 1e-10, worst


@pytest.mark.parametrize("tau_phi", [0.01, 0.9667])
def test_reduced_5x5_riccati_matches_scipy(host, tau_phi):
    rng = np.random.default_rng(1)
    q, r = [1., 1., 0.1, 0.01, 0.01], [8., 1.]
    q5, pq = _arr(q); r2, pr = _arr(r)
    worst = 0.
    for _ in range(200):
        v, phi = rng.uniform(2., 40.), rng.uniform(-1.2, 1.2)
        st, pst = _arr(np.zeros(7)); Kp, pK = _arr(np.zeros(10))
        assert host.d2dx_host_lqr5_gain(pq, pr, v, phi, tau_phi, 1., pst, pK) == 1
        Ks = scipy_gain5(v, phi, q, r, tau_phi, 1.)
        worst = max(worst, np.abs(Kp.reshape(2, 5) - Ks).max() / np.abs(Ks).max())
        # warm restart from the converged state at a nearby reference
        assert host.d2dx_host_lqr5_gain(pq, pr, v * 1.01, phi + 0.01, tau_phi, 1., pst, pK) == 1
        Ks = scipy_gain5(v * 1.01, phi + 0.01, q, r, tau_phi, 1.)
        worst = max(worst, np.abs(Kp.reshape(2, 5) - Ks).max() / np.abs(Ks).max())
    assert worst < 1e-10, worst
