"""Spline references (TrajSpline, SplineOne; d2d/trajectory_factory.py:189-221) against golden vectors of the
unmodified reference and, for way points the upstream constructor cannot take, against FITPACK evaluated on the host."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "drone-sim-python_b200"))

pytestmark = pytest.mark.gpu


def test_traj_spline_samples_and_closed_loop():
    from d2d_b200 import dynamic as ddyn, guidance as ddg, trajectory_factory as ddtf
    from d2d_b200.simulation import run_simulation
    g = np.load(os.path.join(HERE, "golden", "spline.npz"))
    traj, _ = ddtf.get("spline")
    assert abs(traj.duration - float(g["duration"])) < 1e-12
    Y = traj.get_many(g["ts"])
    scale = np.abs(g["Y"]).max(axis=(0, 2))[None, :, None]
    np.testing.assert_allclose(Y / scale, g["Y"] / scale, rtol=0, atol=1e-11)
    np.testing.assert_allclose(traj.get(5.0), g["Y"][np.argmin(np.abs(g["ts"] - 5.0))], atol=2.0)     # same call shape (4, 2)
    time, X0 = g["time"], g["X0"]
    ctl = ddg.DFFFController(traj, ddyn.Aircraft(), ddg.WindField(list(g["wind"])))
    X, U, _ = run_simulation(time, ddyn.Aircraft(), ddg.WindField(list(g["wind"])), ctl, X0, np.zeros((len(time), 5)))
    np.testing.assert_allclose(X, g["X"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(U, g["U"], rtol=0, atol=1e-8)


def test_spline_one_and_custom_waypoints():
    import scipy.interpolate as interpolate
    from d2d_b200 import trajectory_factory as ddtf
    g = np.load(os.path.join(HERE, "golden", "spline.npz"))
    one = ddtf.SplineOne(g["one_xs"], g["one_ys"])
    Y1 = np.array([one.get(t) for t in g["one_t"][::7]])
    np.testing.assert_allclose(Y1, g["one_Y"][::7], rtol=0, atol=1e-10)
    # eight way points -> interior knots (the upstream constructor only works with its default way points)
    wp = np.array([[0, 0], [30, 10], [60, -5], [90, 20], [120, 0], [150, 30], [180, 10], [200, 0.]])
    tr = ddtf.TrajSpline(waypoints=wp, duration=40.)
    assert len(tr.segments()) == 4
    lam = np.linspace(0, 40., len(wp))
    spl = [interpolate.InterpolatedUnivariateSpline(lam, wp[:, i], k=4) for i in range(2)]
    ts = np.linspace(0., 95., 333)
    ref = np.array([np.array([s.derivatives(np.fmod(t, 40.)) for s in spl])[:, :4].T for t in ts])
    Y = tr.get_many(ts)
    scale = np.abs(ref).max(axis=(0, 2))[None, :, None]
    np.testing.assert_allclose(Y / scale, ref / scale, rtol=0, atol=1e-11)
