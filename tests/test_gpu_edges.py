"""Edge cases: ragged batch sizes, single-sample grids, perturbation events, degenerate formations / problems, and
the error behaviour of the C ABI."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _pop(B, seed=1):
    rng = np.random.default_rng(seed)
    cx, cy, r, v, a0 = rng.uniform(-50, 50, B), rng.uniform(-50, 50, B), rng.uniform(30, 60, B), rng.uniform(10, 12, B), rng.uniform(0, 6, B)
    X0 = np.stack([cx + r * np.cos(a0) + 2., cy + r * np.sin(a0) - 1., a0 + np.pi / 2, 0 * a0, v], 1)
    return (cx, cy, r, v, a0), X0, rng.normal(0, 1., (B, 2))


@pytest.mark.parametrize("B", [1, 31, 127, 129, 1000])
def test_ragged_batch_sizes_match_single_runs(B):
    """Scenario b of a batch of any size (TMA-staged full tiles, gather-staged ragged tails) equals the same scenario run alone."""
    from d2d_b200 import simulation, trajectory
    p, X0, wind = _pop(B)
    time = np.arange(0, 1.0, 0.01)
    res = simulation.rollout(time, trajectory.CircleBatch(*p), wind, X0)
    for b in sorted({0, B // 2, B - 1}):
        one = simulation.rollout(time, trajectory.CircleBatch(*[a[b:b + 1] for a in p]), wind[b:b + 1], X0[b:b + 1])
        np.testing.assert_array_equal(res.X[b], one.X[0])
        np.testing.assert_array_equal(res.U[b], one.U[0])
    assert res.X.shape == (B, 100, 5) and not res.flags.any()
    np.testing.assert_allclose(res.pop_sum_sq_err, res.sum_sq_err.sum(), rtol=1e-12)
    assert res.pop_max_err == res.max_err.max()


def test_single_sample_grid_and_mixed_trajectory_types():
    from d2d_b200 import simulation, trajectory, trajectory_factory as ddtf
    from oracle import d2d_oracle as orc
    tr = [trajectory.TrajectoryCircle(), ddtf.TrajSquare(), ddtf.TrajMinSnapDemo(), trajectory.TrajectoryLine([0, 0], [10, 5])]
    X0 = np.array([[60., 30, 1.5, 0, 10], [0, 1, 0, 0, 10], [0, 0, 0.1, 0, 10], [0, 0, 0.4, 0, 9]])
    one = simulation.rollout(np.array([0.3]), tr, [1., -1.], X0)                  # T = 1: only the controller evaluation
    assert one.X.shape == (4, 1, 5)
    np.testing.assert_array_equal(one.X[:, 0], X0)
    otr = [orc.Circle(), orc.traj_square(), orc.traj_minsnap_demo(), orc.Line([0, 0], [10, 5])]
    for b in range(4):
        U, _, _ = orc.dfff_control(otr[b], X0[b].copy(), 0.3, [1., -1.])
        np.testing.assert_allclose(one.U[b, 0], U, rtol=0, atol=1e-11)
    res = simulation.rollout(np.arange(0, 3., 0.01), tr, [1., -1.], X0)           # mixed types -> generic kernel
    for b in range(4):
        Xo, Uo, _, _, _ = orc.run_simulation(np.arange(0, 3., 0.01), otr[b], [1., -1.], X0[b])
        np.testing.assert_allclose(res.X[b], Xo, rtol=0, atol=1e-9)


def test_perturbation_events_sparse_and_dense():
    """`X[i] += perts[i]` (05_test_simulation.py:32): several events on one scenario, none on its neighbours."""
    from d2d_b200 import simulation, trajectory
    from oracle import d2d_oracle as orc
    time = np.arange(0, 2., 0.01)
    T = len(time)
    perts = [None, np.zeros((T, 5)), None]
    perts[1][50] = [1., -2., 0.3, 0., 0.5]; perts[1][51, 1] = 4.; perts[1][T - 1, 0] = -3.; perts[1][0, 0] = 99.   # perts[0] is never applied
    tr = [trajectory.TrajectoryCircle(alpha0=k) for k in range(3)]
    X0 = np.array([[60., 30 + k, 1.5, 0, 10] for k in range(3)])
    res = simulation.rollout(time, tr, [0., 2.], X0, perts=perts)
    for b in range(3):
        Xo, Uo, _, _, _ = orc.run_simulation(time, orc.Circle(alpha0=b), [0., 2.], X0[b], perts[b])
        np.testing.assert_allclose(res.X[b], Xo, rtol=0, atol=1e-9)
        np.testing.assert_allclose(res.U[b], Uo, rtol=0, atol=1e-9)
    chunked = simulation.rollout(time, tr, [0., 2.], X0, perts=perts, chunk_steps=51)    # an event on a chunk boundary
    np.testing.assert_allclose(chunked.X, res.X, rtol=0, atol=1e-13)


def test_degenerate_formations_and_problems():
    from d2d_b200 import simulation
    from d2d_b200.collocation import CollocationProblem, CostSpec
    from oracle import d2d_oracle as orc
    X1 = np.array([20, 30, -np.pi / 2, 0, 10.])
    one = simulation.formation_rollout(np.zeros((3, 1, 2)), 60., 1, 50, 0.05, 4e-4, 15, 20, np.zeros(0), X1)   # n_ac = 1: pure GVF
    Xo, Uo, *_ = orc.run_formation(np.zeros((1, 2)), 60, 1, 2.5, 4e-4, 15, 20, np.zeros(0), nsub=5)
    np.testing.assert_allclose(one["X"][2], Xo, rtol=0, atol=1e-9)
    big = simulation.formation_rollout(np.zeros((2, 32, 2)), 60., 32, 40, 0.05, 4e-4, 15, 20, np.ones(31) * 2 * np.pi / 32, X1)   # one formation per warp
    Xo, *_ = orc.run_formation(np.zeros((32, 2)), 60, 32, 2.0, 4e-4, 15, 20, np.ones(31) * 2 * np.pi / 32, nsub=5)
    np.testing.assert_allclose(big["X"][1], Xo, rtol=0, atol=1e-9)
    rng = np.random.default_rng(0)
    for n_ac, N in ((1, 2), (2, 3), (5, 33), (3, 129)):                            # N = 2: a single defect per equation
        free = rng.normal(0, 2., 5 * n_ac * N); free[4 * n_ac * N:] += 10.
        p = CollocationProblem(n_ac, N, 0.1, wind=(1., 0.), cost=CostSpec(vsp=10., kvel=1., kbank=1., kcol=2., rcol=5., all_pairs=True))
        res, jac, cost, grad = p.evaluate(free)
        np.testing.assert_allclose(res, orc.colloc_residual(free, N, n_ac, 0.1, (1., 0.), []), rtol=1e-10, atol=1e-11)
        np.testing.assert_allclose(jac, orc.colloc_jac_compact(free, N, n_ac, 0.1).reshape(-1), rtol=1e-10)
        co, go = orc.cost_and_grad(free, N, n_ac, dict(vsp=10., kvel=1., kbank=1., kcol=2., rcol=5., pairs="all"), multi=n_ac > 1)
        np.testing.assert_allclose(cost, co, rtol=1e-10); np.testing.assert_allclose(grad, go, rtol=1e-10, atol=1e-13)


def test_c_abi_error_behaviour():
    """Invalid arguments return an error code and a message; nothing is launched, nothing throws."""
    import d2d_b200
    from d2d_b200 import _lib
    eng = d2d_b200.get_engine()
    lib = _lib.lib
    assert lib.d2dx_rollout_dfff(eng.h, None, None, 0, 1, 1, 0, None, None, None) == 1
    assert b"null" in lib.d2dx_last_error()
    s, o = _lib.Scenarios(), _lib.RolloutOut()
    s.B = 0
    assert lib.d2dx_rollout_dfff(eng.h, C.byref(s), C.c_void_p(8), 0, 1, 1, 0, None, C.byref(o), None) == 1
    p = _lib.CollocProblem(); p.n_ac, p.N, p.h, p.in_div = 1, 1, 0.1, 1
    assert lib.d2dx_colloc_eval(eng.h, C.byref(p), 1, C.c_void_p(8), 0, 1, C.c_void_p(8), None, None, None, None, None) == 1
    assert b"N=1" in lib.d2dx_last_error()
    f, fo = _lib.Formations(), _lib.FormationOut()
    f.F, f.n_ac = 1, 33
    assert lib.d2dx_rollout_formation(eng.h, C.byref(f), 0.05, 0, 1, 1, C.byref(fo), None) == 1
    h2 = C.c_void_p()
    assert lib.d2dx_create(999, C.byref(h2)) != 0
    with pytest.raises(ValueError):
        from d2d_b200 import simulation, trajectory
        simulation.rollout(np.arange(3) * 0.01, [trajectory.TrajectoryCircle()], [0., 0.], np.zeros((2, 5)))
