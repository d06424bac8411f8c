"""5-state LQR tracker on sampled references through the C ABI vs the unmodified reference
(Controllers.py + implement_controller of 10_opt_traj_tracking.py under fixed-step RK4, nsub = 10)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-9


@pytest.mark.parametrize("tag", ["simple", "opt"])
def test_tracker_against_reference_golden(golden, tag):
    from d2d_b200 import controllers
    g = golden["tracker"]
    X, U, Xr, Yd, Ydd, dX, K, flags = controllers.track(g[f"{tag}/time"], g[f"{tag}/x_ref"], g[f"{tag}/y_ref"], g[f"{tag}/wind"],
                                                        g[f"{tag}/X0s"], return_gain=True)
    assert not flags.any()
    np.testing.assert_allclose(X, g[f"{tag}/X"], rtol=0, atol=TOL)
    np.testing.assert_allclose(U, g[f"{tag}/U"], rtol=0, atol=TOL)
    np.testing.assert_allclose(Xr, g[f"{tag}/Xr"], rtol=0, atol=TOL)
    np.testing.assert_allclose(dX, g[f"{tag}/dX"], rtol=0, atol=TOL)
    np.testing.assert_allclose(Yd, g[f"{tag}/Yd"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(Ydd, g[f"{tag}/Ydd"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(K[:-1], g[f"{tag}/K"], rtol=1e-9, atol=1e-9)
    print(tag, "max |dX| =", np.abs(X - g[f"{tag}/X"]).max(), " max |dK| =", np.abs(K[:-1] - g[f"{tag}/K"]).max())


@pytest.mark.parametrize("tag", ["hf", "stline", "inf"])
def test_tracker_on_the_other_shipped_planner_outputs(golden, tag):
    """opt_states_hf.csv, opt_states_st_line.csv and inf_traj_10s.csv (phase 3 of the full-mission scripts)."""
    from d2d_b200 import controllers
    g = golden["tracker"]
    X, U, *_rest = controllers.track(g[f"{tag}/time"], g[f"{tag}/x_ref"], g[f"{tag}/y_ref"], g[f"{tag}/wind"], g[f"{tag}/X0s"])
    assert not _rest[-1].any()
    np.testing.assert_allclose(X, g[f"{tag}/X"], rtol=0, atol=TOL)
    np.testing.assert_allclose(U, g[f"{tag}/U"], rtol=0, atol=TOL)


def test_tracker_drop_in_call_sequence(golden):
    """implement_controller(n_ac, df, v, w, X0s) with a DataFrame, and the single-call classes."""
    import pandas as pd
    from d2d_b200 import controllers, dynamic
    g = golden["tracker"]
    cols = {"time": g["simple/time"]}
    for i in range(4):
        cols[f"x_{i + 1}"], cols[f"y_{i + 1}"], cols[f"psi_{i + 1}"] = g["simple/x_ref"][:, i], g["simple/y_ref"][:, i], 0 * g["simple/time"]
    X, U, Xr, Yd, Ydd, t, dX = controllers.implement_controller(4, pd.DataFrame(cols), 10, [0, 0], g["simple/X0s"])
    np.testing.assert_allclose(X, g["simple/X"], rtol=0, atol=TOL)
    np.testing.assert_allclose(U, g["simple/U"], rtol=0, atol=TOL)
    # ComputeFlatness with a non-zero third derivative
    for i in range(len(g["flat/Y"])):
        Y = g["flat/Y"][i]
        Xr1, Ur1 = controllers.DiffFlatness(list(g["flat/W"][i])).ComputeFlatness(0., Y[0], Y[1], Y[2], Y[3])
        np.testing.assert_allclose(Xr1, g["flat/Xr"][i], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(Ur1, g["flat/Ur"][i], rtol=1e-11, atol=1e-12)
    # ComputeGain step by step reproduces the first steps of the loop
    ctrl, ac = controllers.DiffController([0, 0]), dynamic.Aircraft()
    Fdx, Fdy, Fddx, Fddy = controllers.ComputeDerivatives(g["simple/x_ref"][:, 0], g["simple/y_ref"][:, 0], 0.1)
    for i in (1, 2, 3):
        Xr_, dX_, U_ = ctrl.ComputeGain(0., g["simple/X"][i - 1, 0], [g["simple/x_ref"][i, 0], g["simple/y_ref"][i, 0]], [Fdx[i], Fdy[i]],
                                        [Fddx[i], Fddy[i]], [0, 0], ac)
        np.testing.assert_allclose(U_, g["simple/U"][i - 1, 0], rtol=0, atol=TOL)
        np.testing.assert_allclose(ctrl.K[-1], g["simple/K"][i - 1, 0], rtol=1e-9, atol=1e-9)


def test_lqr5_gain_over_wide_range_against_scipy():
    """The in-kernel 6-unknown Riccati solve vs scipy's 5x5 CARE (through the oracle), cold and far from the envelope."""
    from oracle import d2d_oracle as orc
    import d2d_b200
    eng = d2d_b200.get_engine()
    rng = np.random.default_rng(8)
    n = 200
    Ys = rng.normal(0, 1, (n, 4, 2)) * np.array([50., 8., 2., 0.])[None, :, None]
    Ys[:, 1] += rng.choice([-1, 1], (n, 1)) * rng.uniform(3, 25, (n, 1))
    W = rng.normal(0, 1.5, (n, 2)); X = rng.normal(0, 1, (n, 5))
    for tau_phi in (0.01, 0.9667):
        ac = np.tile([[tau_phi], [1.]], (1, n))
        U, Xr, dX, K = eng.tracker_control(eng.to_device(X.T.copy()), eng.to_device(Ys.reshape(n, 8).T.copy()), eng.to_device(W.T.copy()), eng.to_device(ac))
        K = K.cpu().numpy().T.reshape(n, 2, 5)
        worst = 0.
        for i in range(n):
            _, _, _, Ko = orc.tracker_gain(X[i].copy(), Ys[i, 0], Ys[i, 1], Ys[i, 2], Ys[i, 3], W[i], tau_phi, 1.)
            worst = max(worst, np.abs(K[i] - Ko).max() / np.abs(Ko).max())
        print("tau_phi", tau_phi, "worst relative gain error", worst)
        assert worst < 1e-9
