"""Host-side pieces of the planner NLP solve that need no GPU: the oracle's shooting restatement against the oracle's
collocation residual, and the batched L-BFGS driver on analytic problems."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "drone-sim-python_b200"))
sys.path.insert(0, os.path.join(HERE, ".."))

from oracle import d2d_oracle as orc  # noqa: E402


def test_oracle_shooting_point_zeroes_the_collocation_defects():
    rng = np.random.default_rng(1)
    n_ac, N, h, wind = 3, 50, 0.05, (1.5, -0.7)
    phi, v = rng.uniform(-0.5, 0.5, (n_ac, N)), rng.uniform(9, 15, (n_ac, N))
    p0 = rng.uniform(-10, 10, (3, n_ac))
    free = orc.shoot_free(phi, v, p0, h, wind)
    inst = [(3 * a + k, 0, p0[k, a]) for a in range(n_ac) for k in range(3)]
    res = orc.colloc_residual(free, N, n_ac, h, wind, inst)
    assert np.abs(res).max() < 1e-11
    cost, c, L = orc.shoot_lagrangian(phi, v, p0, p0 + 1.0, h, wind, dict(vsp=12., kvel=1.), np.ones((3, n_ac)), 4.0, multi=True)
    x, y, psi = orc.shoot_states(phi, v, p0, h, wind)
    np.testing.assert_allclose(c, np.stack([x[:, -1], y[:, -1], psi[:, -1]]) - (p0 + 1.0))
    assert abs(L - (cost + c.sum() + 2.0 * (c * c).sum())) < 1e-12


def test_batched_lbfgs_independent_columns():
    """Each column is its own problem: a convex quadratic, Rosenbrock chains with different starts, and one column that
    starts at its optimum (must stay put while the others iterate)."""
    from d2d_b200.shooting import lbfgs
    n, P = 10, 4
    A = torch.diag(torch.logspace(0, 3, n, dtype=torch.float64))
    b = torch.arange(1, n + 1, dtype=torch.float64)

    def fun(x):
        f, g = torch.zeros(P, dtype=torch.float64), torch.zeros_like(x)
        q = x[:, 0]
        f[0], g[:, 0] = 0.5 * q @ A @ q - b @ q, A @ q - b
        for p in (1, 2, 3):
            z = x[:, p]
            f[p] = (100 * (z[1:] - z[:-1] ** 2) ** 2 + (1 - z[:-1]) ** 2).sum()
            g[:-1, p] = -400 * z[:-1] * (z[1:] - z[:-1] ** 2) - 2 * (1 - z[:-1])
            g[1:, p] += 200 * (z[1:] - z[:-1] ** 2)
        return f, g

    x0 = torch.zeros(n, P, dtype=torch.float64)
    x0[:, 2] = -1.2
    x0[:, 3] = 1.0                                            # already optimal
    x, f, g, it = lbfgs(fun, x0, maxit=2000, gtol=1e-9, ftol=0.0)
    np.testing.assert_allclose(x[:, 0].numpy(), (b / torch.diag(A)).numpy(), rtol=1e-7)
    np.testing.assert_allclose(x[:, 1].numpy(), 1.0, atol=1e-6)
    np.testing.assert_allclose(x[:, 2].numpy(), 1.0, atol=1e-6)
    assert torch.equal(x[:, 3], x0[:, 3]) and g.abs().max() < 1e-8


def test_mission_host_logic():
    """Stop rules and the symmetric extension of 11_full_sim_case1.py / 12_full_sim_case2.py (pure host code)."""
    from d2d_b200 import mission
    T, n_ac = 12, 2
    X = np.zeros((T, n_ac, 5)); X[:, :, 0] = 100.
    X0f = np.zeros((n_ac, 5))
    assert mission.first_stop_index(X, X0f) is None
    X[7:, 0, 0] = 1.0; X[5:, 1, 0] = 2.9                      # both inside (3, 3, 0.5 deg) from row 7 on
    assert mission.first_stop_index(X, X0f) == 8               # seen at the top of iteration 8 -> rows [:8]
    X[0, :, 0] = 0.                                            # row 0 (initial state) is never tested by the loop
    assert mission.first_stop_index(X, X0f) == 8
    X[7:, 0, 2] = np.deg2rad(0.6)                              # heading outside 0.5 deg
    assert mission.first_stop_index(X, X0f) is None
    eth = np.full((T, 1), 5.0); eth[0] = 0.; eth[4:] = 0.4
    assert mission.first_stop_index(None, e_theta=eth) == 4     # inside iteration 4 -> rows [:5]
    # symmetric extension: aircraft 0 ends where aircraft 1 starts and vice versa
    t = np.arange(3) * 0.1
    x = np.array([[0., 10.], [5., 5.], [10., 0.]]); y = np.array([[1., -1.], [2., -2.], [-1., 1.]])
    t2, x2, y2, psi2 = mission.ExtendTraj_symm(2, x, y, np.zeros_like(x), t)
    assert len(t2) == 6 and x2.shape == (6, 2)
    np.testing.assert_array_equal(x2[3:, 0], x[:, 1]); np.testing.assert_array_equal(y2[3:, 1], y[:, 0])
    np.testing.assert_allclose(t2[3:], t + t[-1])
    assert mission.ConstructBMatrix(3).tolist() == [[-1, 0], [1, -1], [0, 1]]


def test_experiment_table_matches_upstream_values():
    from d2d_b200 import optyplan_scenarios as S
    assert len(S.scens) == 15 and S.exp_0.t1 == 15.                # building exp_6 sets exp_0.t1 as upstream's class body does
    S.exp_0.t1 = 10.
    assert S.exp_5.cost.spec().obstacles[0] == (0., 20., 10.) and len(S.exp_5.obstacles) == 12
    S.exp_0_2.set_case(3); assert S.exp_0.wind.w == [5., 0.]; S.exp_0_2.set_case(0)
    S.exp_6.set_case(2); assert S.exp_0.p0 == (10, 10, np.pi / 2, 0., 10.) and S.exp_6.label(2) == "2"
    S.exp_0.p0, S.exp_0.p1 = (0., 0., 0., 0., 10.), (0., 30., np.pi, 0., 10.)
    assert S.desc_one(4).startswith("exp_1 combined phi/vel objective\ninitial state 0.0 (0.0, 0.0, 0.0, 0.0, 12.0)")


def test_multi_experiment_table():
    from d2d_b200 import multiopty_scenarios as S
    assert len(S.scens) == 15 and S.get_scen(13) is S.gvf_trial_3ac and "exp_2 4 aicraft" in S.desc_all_scens()
    assert abs(S.exp_2.d - 22.0) < 1e-12 and S.exp_2.p0s[2] == (0., 22.0, -np.pi / 2, 0., 12.)
    S.exp_5.set_case(1); assert S.exp_5.cost.kcol == 10. and S.exp_5.cost.rcol == 10. and S.exp_5.label(1) == "obj AntiCol"
    S.exp_5.set_case(0); assert np.isnan(S.exp_5.cost.kcol) and S.exp_5.cost.rcol == 3.
    S.exp_4_1.set_case(2); assert S.exp_4_1.obstacles == ((70, -10, 15), (70, 25, 15)) and S.exp_4_1.cost.vsp == 14.
    S.exp_3_1.set_case(1); assert S.exp_3_1.cost.c == (25, -10)
    assert S.trap_4.cost.kvel == 70. and S.trap_4.x_constraint == (-150, 150)


def test_oracle_adjoint_gradient_against_finite_differences():
    rng = np.random.default_rng(4)
    n_ac, N, h, wind = 2, 40, 0.1, (0.7, -0.4)
    phi, v = rng.uniform(-0.4, 0.4, (n_ac, N)), rng.uniform(9.5, 14.5, (n_ac, N))
    p0, p1 = rng.uniform(-5, 5, (3, n_ac)), rng.uniform(0, 30, (3, n_ac))
    spec, lam, rho = dict(vsp=12., kvel=3., kbank=2., obj_scale=1.5), rng.normal(0, 1, (3, n_ac)), 4.0
    L, dphi, dv, cost, c = orc.shoot_value_and_grad(phi, v, p0, p1, h, wind, spec, lam, rho, multi=True)
    co, cc, Lo = orc.shoot_lagrangian(phi, v, p0, p1, h, wind, spec, lam, rho, multi=True)
    assert abs(L - Lo) < 1e-12 * abs(Lo) and abs(cost - co) < 1e-14 and np.abs(c - cc).max() == 0
    for (k, a, i) in ((0, 0, 1), (0, 1, 39), (1, 0, 17), (1, 1, 0), (0, 0, 0), (1, 1, 39)):
        d = np.zeros((2, n_ac, N)); d[k, a, i] = 1e-4
        F = lambda s_: orc.shoot_lagrangian(phi + s_ * d[0], v + s_ * d[1], p0, p1, h, wind, spec, lam, rho, multi=True)[2]
        fd = (8 * (F(1) - F(-1)) - (F(2) - F(-2))) / 12e-4
        an = (dphi if k == 0 else dv)[a, i]
        assert abs(an - fd) <= 1e-7 * max(1., abs(fd)) + 2e-11 * abs(Lo), (k, a, i, an, fd)


def test_inputs_from_path_on_a_circle():
    """A circle of radius 40 m flown at 12 m/s needs tan(phi) = v^2 / (g r)."""
    from d2d_b200.planner import inputs_from_path
    h, r, v = 0.1, 40., 12.
    a = np.arange(200) * h * v / r
    phi, vv = inputs_from_path(r * np.cos(a), r * np.sin(a), h, (-0.7, 0.7), (9., 15.))
    np.testing.assert_allclose(vv[5:-5], v, rtol=1e-3)
    np.testing.assert_allclose(np.tan(phi[20:-20]), v * v / (9.81 * r), rtol=2e-3)


def test_opty_input_block_order_mapping():
    """MultiPlanner._to_opty_order: numeric input blocks [phi_0..phi_{n-1} | v_0..v_{n-1}] -> opty's name-sorted blocks
    (phi0, phi1, phi10, phi11, phi2, ... then v0, v1, v10, ...), states untouched."""
    from d2d_b200 import planner as pl, multiopty_scenarios as S

    class scen12(S.exp_0):
        p0s = tuple((float(i), 0., 0., 0., 10.) for i in range(12))
        p1s = tuple((float(i), 50., 0., 0., 10.) for i in range(12))
        hz, t1 = 1., 3.
    p = pl.MultiPlanner(scen12, initialize=False)
    n, N = 12, p.num_nodes
    free = np.arange(5 * n * N, dtype=float)[None]
    out = p._to_opty_order(free)
    np.testing.assert_array_equal(out[:, :3 * n * N], free[:, :3 * n * N])
    names = sorted([f"phi{i}" for i in range(n)] + [f"v{i}" for i in range(n)])
    o = 3 * n * N
    for rank, nm in enumerate(names):
        i = int(nm.lstrip("phiv"))
        src = o + (i if nm.startswith("phi") else n + i) * N
        np.testing.assert_array_equal(out[0, o + rank * N:o + (rank + 1) * N], free[0, src:src + N])
    assert names[:4] == ["phi0", "phi1", "phi10", "phi11"]
