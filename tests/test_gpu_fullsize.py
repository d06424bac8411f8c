"""BASELINE.json's full sizes: the C5 Monte-Carlo sweep (10^6 scenarios x 10^4 RK4 steps) checked on a random sample
against the C oracle and through size-independent properties; the batched collocation at 4096 x C3."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_c5_full_size_sample_against_c_oracle():
    import bench
    from d2d_b200 import _lib
    from d2d_b200.simulation import MonteCarloRollout
    from oracle import c_oracle as co
    B, T = 10 ** 6, 10 ** 4
    w = bench.workload(B, 12345)
    X0 = bench.flat_state0(w) + w["noise"]
    time = np.arange(T + 1) * 0.01
    mc = MonteCarloRollout(B, time, _lib.SEG_CIRCLE, log_every=100, n_chunks=10, host_log=False)
    par = np.zeros((6, B)); par[1], par[2], par[3], par[4], par[5] = w["cx"], w["cy"], w["r"], w["v"] / w["r"], w["a0"]
    mc.set_inputs(par, w["wind"], X0)
    out = mc.run()
    assert not out["flags"].any()
    Xlog = mc.d_Xlog.cpu().numpy()                       # [101][5][B]
    # chunk boundaries are invisible: row 0 is X0, the last row is the final state
    np.testing.assert_array_equal(Xlog[0].T, X0)
    np.testing.assert_array_equal(Xlog[-1].T, out["X_final"])
    # population reductions = reductions of the per-scenario outputs (atomics vs host sum)
    np.testing.assert_allclose(out["pop_sum_sq_err"], out["sum_sq_err"].sum(), rtol=1e-9)
    assert out["pop_max_err"] == out["max_err"].max()
    # a random sample through the C oracle (generic CARE, plain RK4), every logged sample
    rng = np.random.default_rng(7)
    idx = np.sort(rng.choice(B, 384, replace=False))
    ty, opar = co.circle_par(w["cx"][idx], w["cy"][idx], w["r"][idx], w["v"][idx], w["a0"][idx])
    ref = co.rollout(time, ty, opar, w["wind"][idx], X0[idx], log_every=100)
    assert ref["failed"] == 0
    tracked = ref["max_err"] < 10.0                      # scenarios the controller can follow (bank <= 45 deg); the rest diverge chaotically
    assert tracked.sum() > 300
    got = Xlog[:, :, idx].transpose(2, 0, 1)             # (n, 101, 5)
    err = np.abs(got[tracked] - ref["X"][tracked]).max()
    print("C5 full size: max |dX| over", int(tracked.sum()), "tracked sample scenarios x 101 logged samples =", err)
    assert err < 1e-9
    np.testing.assert_allclose(out["sum_sq_err"][idx][tracked], ref["sum_sq_err"][tracked], rtol=1e-8)
    np.testing.assert_allclose(out["max_err"][idx][tracked], ref["max_err"][tracked], rtol=1e-8)
    # untracked ones: same qualitative outcome (they left the reference in both)
    assert (out["max_err"][idx][~tracked] > 9.0).all()


def test_c3_batch_4096_properties(golden):
    """4096 C3 problems in one launch: each row equals its single-problem evaluation; dense and compact Jacobians agree;
    the exact-gradient mode is the finite-difference derivative of the cost."""
    from d2d_b200.collocation import CollocationProblem, CostSpec
    g = golden["colloc"]
    N, h, n_prob = 1001, 0.02, 4096
    rng = np.random.default_rng(12345)
    sig = np.repeat([1., 1., 0.1, 0.05, 0.5], N)
    free = g["c3/sol"][None] + rng.normal(0, 1., (n_prob, 5 * N)) * sig
    inst = [(int(k), int(n), v) for (k, n, v) in g["c3/inst"]]
    cs = CostSpec(vsp=12., kvel=1., kbank=2., kobs=0.5, obstacles=[(5, 15, 10.)], obs_kind=1, exact_grad=True)
    pc = CollocationProblem(1, N, h, inst=inst, cost=cs, layout="compact")
    pd = CollocationProblem(1, N, h, inst=inst, cost=cs, layout="dense")
    res, jac, cost, grad = pc.evaluate(free)
    for p in (0, 1234, n_prob - 1):
        r1, j1, c1, g1 = pc.evaluate(free[p])
        np.testing.assert_array_equal(res[p], r1); np.testing.assert_array_equal(jac[p], j1)
        np.testing.assert_array_equal(grad[p], g1); assert cost[p] == c1
    jd = pd.con_jac(free[:8])
    rc, cc = pc.jacobianstructure(); rd, cd = pd.jacobianstructure()
    for p in range(8):
        A = np.zeros((pc.num_constraints, pc.num_free)); A[rc, cc] = jac[p]
        Bm = np.zeros_like(A); np.add.at(Bm, (rd, cd), jd[p])
        np.testing.assert_array_equal(A, Bm)
    # Jacobian = derivative of the residual, gradient = derivative of the cost (central differences on one problem)
    f0 = free[7]
    for k in rng.choice(5 * N, 6, replace=False):
        d = np.zeros(5 * N); d[k] = 1e-6
        fd_c = (pc.obj(f0 + d) - pc.obj(f0 - d)) / 2e-6
        assert abs(fd_c - grad[7][k]) < 1e-6 * max(1., abs(fd_c))
        fd_r = (pc.con(f0 + d) - pc.con(f0 - d)) / 2e-6
        col = np.zeros(pc.num_constraints); sel = cc == k; col[rc[sel]] = jac[7][sel]
        np.testing.assert_allclose(col, fd_r, rtol=0, atol=2e-5)


def test_long_horizon_against_c_oracle():
    """10^5 control steps (1000 s of flight, circle phases up to ~500 rad): no drift between the engine and the C oracle."""
    from d2d_b200 import simulation, trajectory
    from oracle import c_oracle as co
    import bench
    B, T = 64, 10 ** 5
    w = bench.workload(B, 777)
    w["r"] = np.clip(w["r"], 35., None)                 # keep the required bank below the 45 deg limit: trackable circles
    X0 = bench.flat_state0(w) + w["noise"]
    time = np.arange(T + 1) * 0.01
    res = simulation.rollout(time, trajectory.CircleBatch(w["cx"], w["cy"], w["r"], w["v"], w["a0"]), w["wind"], X0, log_every=1000)
    ty, par = co.circle_par(w["cx"], w["cy"], w["r"], w["v"], w["a0"])
    ref = co.rollout(time, ty, par, w["wind"], X0, log_every=1000)
    ok = ref["max_err"] < 10.
    assert ok.sum() >= B // 2 and not res.flags.any()
    err = np.abs(res.X[ok] - ref["X"][ok]).max()
    print("long horizon: max |dX| =", err, "over", int(ok.sum()), "scenarios x 101 logged samples")
    assert err < 1e-9
