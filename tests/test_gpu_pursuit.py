"""Pure-pursuit controller and its closed loop (d2d/guidance.py:204-245) against golden vectors of the
unmodified reference (tests/golden/pursuit.npz: square patrol and a circle with wind)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "drone-sim-python_b200"))

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(HERE, "golden", "pursuit.npz"))


@pytest.mark.parametrize("tag,factory", [("square", "TrajSquare"), ("circle", "TrajCircle")])
def test_pursuit_closed_loop_against_reference(g, tag, factory):
    from d2d_b200 import dynamic as ddyn, guidance as ddg, trajectory_factory as ddtf
    from d2d_b200.simulation import run_simulation, pursuit_rollout
    traj = getattr(ddtf, factory)()
    ctl = ddg.PurePursuitControler(traj)
    np.testing.assert_allclose(ctl.pts_2d, g[f"{tag}/pts"], rtol=0, atol=1e-12)           # the sampled path itself
    time, X0, w = g[f"{tag}/time"], g[f"{tag}/X0"], g[f"{tag}/wind"]
    X, U, Yref = run_simulation(time, ddyn.Aircraft(), ddg.WindField(list(w)), ctl, X0, np.zeros((len(time), 5)))
    np.testing.assert_allclose(X, g[f"{tag}/X"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(U, g[f"{tag}/U"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(np.array(ctl.carrot), g[f"{tag}/carrot"], rtol=0, atol=1e-12)
    # single call, and a batch with the same aircraft replicated plus shifted copies
    ctl2 = ddg.PurePursuitControler(traj)
    for k in (0, 100, len(time) - 1):
        np.testing.assert_allclose(ctl2.get(g[f"{tag}/X"][k], time[k]), g[f"{tag}/U"][k], rtol=0, atol=1e-9)
    X0b = np.tile(X0, (67, 1)); X0b[1:, 0] += np.linspace(-3, 3, 66)
    Xb, Ub, idx = pursuit_rollout(ctl, time[:300], X0b, w)
    np.testing.assert_allclose(Xb[:, 0], g[f"{tag}/X"][:300], rtol=0, atol=1e-9)
    np.testing.assert_array_equal(idx[:299, 0], g[f"{tag}/idx"][:299])
    assert np.isfinite(Xb).all() and Xb.shape == (300, 67, 5)


def test_vel_controler_host_arithmetic():
    from d2d_b200.guidance import VelControler
    v = VelControler()
    assert v.get(3.0, 1.0) == 10. - min(2. * 2.0 + 0.001 * 2.0, 4)
    assert v.get(100., 0.) == 6.0 and abs(v.sum_err - 12.0) < 1e-12
