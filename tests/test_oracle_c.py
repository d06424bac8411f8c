"""The C restatement of the closed-loop oracle against the reference's golden vectors and the NumPy oracle."""
import numpy as np

from oracle import c_oracle as co
from oracle import d2d_oracle as orc


def test_c1_against_reference_golden(golden):
    g = golden["dfff_c1"]
    ty, par = co.circle_par([30.], [30.], [30.], [10.], [3 * np.pi / 2])
    r = co.rollout(g["time"], ty, par, g["wind"], g["X0"][None], want_K=True)
    assert r["failed"] == 0
    np.testing.assert_allclose(r["X"][0], g["X"], rtol=0, atol=1e-11)
    np.testing.assert_allclose(r["U"][0], g["U"], rtol=0, atol=1e-11)
    np.testing.assert_allclose(r["K"][0].reshape(-1, 2, 3), g["K"][:, :, :3], rtol=0, atol=1e-11)
    d2 = np.sum(np.square(g["X"][:, :2] - g["Xref"][:, :2]), axis=1)
    np.testing.assert_allclose(r["sum_sq_err"][0], d2.sum(), rtol=1e-10)


def test_lqr_against_scipy_golden(golden):
    u = golden["units"]
    for i in range(len(u["K"])):
        K, rc = co.lqr3(u["A"][i][:3, :3], u["A"][i][:3, 3:])
        assert rc == 0
        np.testing.assert_allclose(K, u["K"][i], rtol=1e-11, atol=1e-12)


def test_line_and_minsnap_against_reference_golden(golden):
    g = golden["dfff_scenarios"]
    # 'line' scenario without its perturbation: first 600 samples
    tr = orc.Line([0, 25], [100, 25], v=10.)
    par = np.zeros((1, 17)); par[0, 1:5] = [0, 25, tr.un[0] * 10., tr.un[1] * 10.]
    time = np.arange(0, 12., 0.01)[:600]
    r = co.rollout(time, [co.T_LINE], par, [0., 0.], np.array([[10., 10, 0, 0, 10]]))
    np.testing.assert_allclose(r["X"][0][::5], g["line/0/X"][:120], rtol=0, atol=1e-10)
    ms = orc.traj_minsnap_demo()
    par = np.zeros((1, 17)); par[0, 1:9] = ms._polys[0].coefs[0]; par[0, 9:17] = ms._polys[1].coefs[0]
    time = np.arange(0., ms.duration, 0.01)
    r = co.rollout(time, [co.T_POLY], par, g["minsnap/0/wind"], g["minsnap/0/X0"][None])
    np.testing.assert_allclose(r["X"][0][::5], g["minsnap/0/X"], rtol=0, atol=1e-10)
    np.testing.assert_allclose(r["U"][0][::5], g["minsnap/0/U"], rtol=0, atol=1e-10)


def test_population_against_numpy_oracle():
    rng = np.random.default_rng(3)
    B, T = 6, 200
    cx, cy, r, v, a0 = rng.uniform(-50, 50, B), rng.uniform(-50, 50, B), rng.uniform(20, 60, B), rng.uniform(10, 15, B), rng.uniform(0, 6.28, B)
    wind = rng.normal(0, 2.5, (B, 2))
    time = np.arange(T) * 0.01
    X0 = np.array([orc.flatness(orc.Circle([cx[b], cy[b]], r[b], v[b], alpha0=a0[b]).get(0.), wind[b])[0] for b in range(B)])
    X0 += rng.normal(0, 1, (B, 5)) * np.array([5, 5, 0.2, 0.05, 0.5])
    ty, par = co.circle_par(cx, cy, r, v, a0)
    out = co.rollout(time, ty, par, wind, X0, nthreads=2)
    for b in range(B):
        Xo, Uo, _, _, _ = orc.run_simulation(time, orc.Circle([cx[b], cy[b]], r[b], v[b], alpha0=a0[b]), wind[b], X0[b])
        np.testing.assert_allclose(out["X"][b], Xo, rtol=0, atol=1e-10)
        np.testing.assert_allclose(out["U"][b], Uo, rtol=0, atol=1e-10)
