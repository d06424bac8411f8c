"""The planner's second-order solver (control-limited DDP, csrc/d2dx_ddp.cuh) on the CPU: csrc/host_check.cu compiles the code of
the GPU thread for the host.  Checked: the result satisfies the reference's collocation constraints (NumPy oracle), reaches the
optimum IPOPT reaches on upstream's experiments (costs of the cached solutions shipped with the reference, recomputed in
tests/golden/make_golden.py -> colloc.npz), keeps the input bounds at every node, and solves a population of random targets."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import ddp_host as dh  # noqa: E402
from oracle import d2d_oracle as orc  # noqa: E402

B30 = (-np.deg2rad(30.), np.deg2rad(30.), 9., 14.)


def _check_feasible(u, xs, N, h, z0, zt, bounds, wind=(0., 0.), tol=1e-7):
    free = np.concatenate([xs[0], xs[1], xs[2], u[0], u[1]])
    inst = [(k, 0, z0[k]) for k in range(3)] + [(k, N - 1, zt[k]) for k in range(3)]
    res = orc.colloc_residual(free, N, 1, h, wind, inst)
    assert np.abs(res).max() < tol, np.abs(res).max()
    assert u[0, 1:].min() >= bounds[0] - 1e-12 and u[0, 1:].max() <= bounds[1] + 1e-12
    assert u[1, 1:].min() >= bounds[2] - 1e-12 and u[1, 1:].max() <= bounds[3] + 1e-12


def test_exp0_reaches_the_ipopt_optimum_and_is_feasible():
    """exp_0 (d2d/optyplan_scenarios.py:9-28: turn around in 10 s, CostAirVel(12), bank <= 30 deg): optimum 0.752912 (the value
    the log-barrier Newton prototype and IPOPT-class solvers reach); bank saturates, so the box QP is exercised."""
    N, h = 101, 0.1
    u, xs, info = dh.solve(N, h, (0., 0.), 12., 1., 0., B30, (0., 0., 0.), (0., 30., np.pi), 0.1, 12.)
    assert info["flag"] == 2 and info["iterations"] < 200
    assert abs(info["cost"] - 0.752912) < 2e-6
    assert np.isclose(np.abs(u[0]).max(), B30[1])            # the bank limit is active
    _check_feasible(u, xs, N, h, (0., 0., 0.), (0., 30., np.pi), B30)
    assert abs(np.mean(np.square(u[1] - 12.)) - info["cost"]) < 1e-12


@pytest.mark.parametrize("t1,ipopt_cost", [(7., 3.908415), (10., 0.758672), (20., 8.1e-8)])
def test_cached_ipopt_experiments(t1, ipopt_cost):
    """exp_0_1 cases on the 50 Hz grid of the cached solutions (src/cache/optyplan_exp0_1_{0,1,3}.npz): cost not above IPOPT's."""
    N, h = int(t1 * 50) + 1, 0.02
    best = None
    for phi0, mode in ((0.1, 0), (0.1, 1), (0.3, 0), (0.3, 1)):      # the planner front end's retry ladder: both regularisations
        u, xs, info = dh.solve(N, h, (0., 0.), 12., 1., 0., B30, (0., 0., 0.), (0., 30., np.pi), phi0, 12., opts=dh.default_options(reg_mode=mode))
        if info["flag"] == 2 and (best is None or info["cost"] < best[2]["cost"]):
            best = (u, xs, info)
    assert best is not None
    assert best[2]["cost"] <= ipopt_cost * (1 + 1e-5) + 1e-9
    _check_feasible(best[0], best[1], N, h, (0., 0., 0.), (0., 30., np.pi), B30)


def test_wind_obstacle_and_bank_cost():
    """wind in the transition (sign of d2d/opty_utils.py:42-43), an obstacle (kind 1) and a bank term: feasible, and moving the
    obstacle onto the straight path raises the cost."""
    N, h = 86, 0.1
    b = (-np.deg2rad(40.), np.deg2rad(40.), 9., 15.)
    kw = dict(obstacles=[(50., -10., 25.)], kobs=0.5, obs_kind=1, obj_scale=1e-2)
    u, xs, info = dh.solve(N, h, (1., 0.5), 12., 0.5, 1., b, (0., 0., 0.), (100., 0., 0.), 0.0, 12., **kw)
    assert info["flag"] == 2
    _check_feasible(u, xs, N, h, (0., 0., 0.), (100., 0., 0.), b, wind=(1., 0.5))
    spec = dict(vsp=12., kvel=0.5, kbank=1., kobs=0.5, obstacles=[(50., -10., 25.)], obs_kind=1, obj_scale=1e-2)
    free = np.concatenate([xs[0], xs[1], xs[2], u[0], u[1]])
    c_o, _ = orc.cost_and_grad(free, N, 1, spec, multi=False)          # the reference's cost classes at the solution
    assert abs(c_o - info["cost"]) < 1e-10 * max(1., abs(c_o))
    u2, xs2, info2 = dh.solve(N, h, (1., 0.5), 12., 0.5, 1., b, (0., 0., 0.), (100., 0., 0.), 0.0, 12., obstacles=[(50., 0., 25.)], kobs=0.5,
                              obs_kind=1, obj_scale=1e-2)
    assert info2["flag"] == 2 and info2["cost"] > info["cost"]


def test_population_of_random_targets():
    """SURVEY 8f #2 / bench `planner_population`: exp_0-grid problems with random terminal targets, every one solved to 1e-8 from
    one start, median sweep count far below the first-order driver's (~1300 ticks)."""
    rng = np.random.default_rng(12345)
    n, N, h = 60, 101, 0.1
    p1 = np.stack([rng.uniform(-10, 10, n), rng.uniform(28, 40, n), np.pi + rng.uniform(-0.5, 0.5, n)], 1)
    its, solved = [], 0
    for p in range(n):
        u, xs, info = dh.solve(N, h, (0., 0.), 12., 1., 0., B30, (0., 0., 0.), p1[p], 0.1, 12.)
        solved += info["flag"] == 2
        its.append(info["iterations"])
        if p % 10 == 0 and info["flag"] == 2:
            _check_feasible(u, xs, N, h, (0., 0., 0.), p1[p], B30)
    assert solved == n and np.median(its) < 150, (solved, np.median(its))
