"""CUDA path (through the C ABI) vs the golden vectors of the unmodified reference and vs the oracle.
Tolerance: north_star's 1e-9 absolute on per-step states under the same fixed-step integrator and dt."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-9


@pytest.fixture(scope="module")
def d2d():
    import d2d_b200
    from d2d_b200 import simulation, trajectory, trajectory_factory, scenario, guidance, dynamic
    d2d_b200.get_engine()
    return d2d_b200


def test_c1_against_reference_golden(d2d, golden):
    from d2d_b200 import simulation, trajectory
    g = golden["dfff_c1"]
    res = simulation.rollout(g["time"], [trajectory.TrajectoryCircle(alpha0=3 * np.pi / 2)], g["wind"], g["X0"][None], log_ref=True)
    assert res.flags[0] == 0
    np.testing.assert_allclose(res.X[0], g["X"], rtol=0, atol=TOL)
    np.testing.assert_allclose(res.U[0], g["U"], rtol=0, atol=TOL)
    np.testing.assert_allclose(res.Xref[0], g["Xref"], rtol=0, atol=TOL)
    np.testing.assert_allclose(res.K[0], g["K"], rtol=0, atol=TOL)
    np.testing.assert_allclose(res.X_final[0], g["X"][-1], rtol=0, atol=TOL)
    # per-scenario reductions = what the log implies
    d2 = np.sum(np.square(g["X"][:, :2] - g["Xref"][:, :2]), axis=1)
    np.testing.assert_allclose(res.sum_sq_err[0], d2.sum(), rtol=1e-9)
    np.testing.assert_allclose(res.max_err[0], np.sqrt(d2.max()), rtol=1e-9)
    print("C1 max |dX| =", np.abs(res.X[0] - g["X"]).max(), " max |dK| =", np.abs(res.K[0] - g["K"]).max())


def test_c1_drop_in_run_simulation(d2d, golden):
    """The reference's own call sequence (05_test_simulation.py:37-53) on the mirrored classes."""
    from d2d_b200 import dynamic, guidance, simulation, trajectory
    g = golden["dfff_c1"]
    traj = trajectory.TrajectoryCircle(alpha0=3 * np.pi / 2)
    ac, wind = dynamic.Aircraft(), guidance.WindField([5, 0])
    ctl = guidance.DFFFController(traj, ac, wind)
    X, U, Yref = simulation.run_simulation(g["time"], ac, wind, ctl, g["X0"], np.zeros((len(g["time"]), 5)))
    np.testing.assert_allclose(X, g["X"], rtol=0, atol=TOL)
    np.testing.assert_allclose(U, g["U"], rtol=0, atol=TOL)
    np.testing.assert_allclose(Yref, g["Yref"], rtol=0, atol=1e-12)
    assert len(ctl.Xref) == len(g["time"]) and np.asarray(ctl.K).shape == (len(g["time"]), 2, 5)


def test_chunked_rollout_equals_single_launch(d2d, golden):
    from d2d_b200 import simulation, trajectory
    g = golden["dfff_c1"]
    tr = [trajectory.TrajectoryCircle(alpha0=3 * np.pi / 2)]
    a = simulation.rollout(g["time"], tr, g["wind"], g["X0"][None])
    b = simulation.rollout(g["time"], tr, g["wind"], g["X0"][None], chunk_steps=137)
    np.testing.assert_allclose(a.X, b.X, rtol=0, atol=1e-13)
    np.testing.assert_allclose(a.sum_sq_err, b.sum_sq_err, rtol=1e-13)
    c = simulation.rollout(g["time"], tr, g["wind"], g["X0"][None], log_every=10)
    np.testing.assert_array_equal(c.X[0], a.X[0][::10])


SCENS = ["line", "line2", "square", "mucir", "mucir2", "patrol", "patrol_2", "patrol_3", "circForm"]


@pytest.mark.parametrize("name", SCENS)
def test_scenario_registry_against_reference_golden(d2d, golden, name):
    """Every runnable scenario of d2d/scenario.py, all aircraft in one launch, full length."""
    from d2d_b200 import scenario, simulation
    g = golden["dfff_scenarios"]
    scen, _ = scenario.get(name)
    assert len(scen.time) == int(g[f"{name}/0/T"])
    Xs, Us, Yrefs = simulation.test_simulation(scen)
    for i in range(len(scen.trajs)):
        np.testing.assert_allclose(Yrefs[i][::25], g[f"{name}/{i}/Yref"], rtol=0, atol=1e-11)
        np.testing.assert_allclose(Xs[i][::5], g[f"{name}/{i}/X"], rtol=0, atol=TOL)
        np.testing.assert_allclose(Us[i][::5], g[f"{name}/{i}/U"], rtol=0, atol=TOL)
        np.testing.assert_allclose(Xs[i][-1], g[f"{name}/{i}/Xlast"], rtol=0, atol=TOL)
        np.testing.assert_allclose(Us[i][-1], g[f"{name}/{i}/Ulast"], rtol=0, atol=TOL)


@pytest.mark.parametrize("name", ["minsnap", "sidemo", "slalom"])
def test_extra_trajectories_against_reference_golden(d2d, golden, name):
    from d2d_b200 import simulation, trajectory_factory as ddtf
    g = golden["dfff_scenarios"]
    traj = {"minsnap": ddtf.TrajMinSnapDemo, "sidemo": ddtf.TrajSiDemo, "slalom": ddtf.TrajSlalom}[name]()
    time = np.arange(0., traj.duration, 0.01)
    assert len(time) == int(g[f"{name}/0/T"])
    np.testing.assert_allclose(traj.get_many(time[::25]), g[f"{name}/0/Yref"], rtol=0, atol=1e-11)
    res = simulation.rollout(time, [traj], g[f"{name}/0/wind"], g[f"{name}/0/X0"][None], log_ref=True)
    np.testing.assert_allclose(res.X[0][::5], g[f"{name}/0/X"], rtol=0, atol=TOL)
    np.testing.assert_allclose(res.U[0][::5], g[f"{name}/0/U"], rtol=0, atol=TOL)
    np.testing.assert_allclose(res.K[0][::25], g[f"{name}/0/K"], rtol=0, atol=TOL)


def test_single_calls_against_reference_golden(d2d, golden):
    """Aircraft.cont_dyn / cont_jac, DiffFlatness, LQR gain, min-snap coefficients."""
    from d2d_b200 import dynamic, guidance, trajectory_factory as ddtf
    u = golden["units"]
    ac = dynamic.Aircraft()
    Xr, Ur, Xd = guidance.DiffFlatness.state_and_input_from_output(u["Ys"], u["Ws"], ac)
    np.testing.assert_allclose(Xr, u["Xr"], rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(Ur, u["Ur"], rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(Xd, u["Xrdot"], rtol=1e-13, atol=1e-13)
    A, B = ac.cont_jac(u["Xr"], u["Ur"], 0., None)
    np.testing.assert_allclose(A, u["A"], rtol=1e-13, atol=1e-14)
    assert B[0][3, 0] == 1 / ac.tau_phi and B[0][4, 1] == 1 / ac.tau_v
    eng = d2d.get_engine()
    Xdot = eng.cont_dyn(eng.to_device(u["Xs"].T.copy()), eng.to_device(u["Us"].T.copy()), eng.to_device(u["Ws"].T.copy()),
                        eng.to_device(np.tile([[0.01], [1.]], (1, len(u["Xs"]))))).cpu().numpy().T
    np.testing.assert_allclose(Xdot, u["Xdot"], rtol=1e-13, atol=1e-13)
    one = ac.cont_dyn(u["Xs"][0], 0., u["Us"][0], guidance.WindField(list(u["Ws"][0])))
    assert isinstance(one, list) and len(one) == 5
    np.testing.assert_allclose(one, u["Xdot"][0], rtol=1e-13)
    ms = ddtf.TrajMinSnapDemo()
    np.testing.assert_array_equal(np.array([p.coefs for p in ms._polys]), u["minsnap_coefs"])
    np.testing.assert_allclose(ms.get(10.0), u["minsnap_get10"], rtol=0, atol=1e-12)


def test_lqr_gain_over_wide_range_against_oracle(d2d):
    """The in-kernel Riccati solve vs scipy's CARE (through the oracle) far outside the flight envelope."""
    from oracle import d2d_oracle as orc
    from d2d_b200 import trajectory
    eng = d2d.get_engine()
    rng = np.random.default_rng(5)
    n = 256
    v = np.exp(rng.uniform(np.log(1.0), np.log(40.), n)); r = rng.uniform(8., 80., n) * rng.choice([-1, 1], n)
    wind = rng.normal(0, 1.5, (n, 2))
    batch = trajectory.CircleBatch(rng.uniform(-50, 50, n), rng.uniform(-50, 50, n), r, v, rng.uniform(0, 6.28, n))
    tab = eng.table(batch.pack())
    X = rng.normal(0, 1, (n, 5)); t = 1.234
    acd = eng.to_device(np.tile([[0.01], [1.]], (1, n)))
    U, Xr, K = eng.dfff_control(tab, eng.to_device(X.T.copy()), t, eng.to_device(wind.T.copy()), acd)
    K = K.cpu().numpy().T.reshape(n, 2, 3); Xr = Xr.cpu().numpy().T
    worst = 0.
    for i in range(n):
        tr = orc.Circle(c=[batch.cx[i], batch.cy[i]], r=r[i], v=v[i], alpha0=batch.alpha0[i])
        _, Xro, Ko = orc.dfff_control(tr, X[i].copy(), t, wind[i])
        np.testing.assert_allclose(Xr[i], Xro, rtol=1e-12, atol=1e-12)
        worst = max(worst, np.abs(K[i] - Ko).max() / max(1., np.abs(Ko).max()))
    assert worst < 1e-10, worst


def test_monte_carlo_population_against_oracle(d2d):
    """A seeded C5-style population (random circles, wind, initial states): 48 scenarios x 300 steps against the
    Python oracle, every step."""
    from oracle import d2d_oracle as orc
    from d2d_b200 import simulation, trajectory
    rng = np.random.default_rng(12345)
    B, T = 48, 301
    cx, cy = rng.uniform(-50, 50, B), rng.uniform(-50, 50, B)
    r, v, a0 = rng.uniform(20, 60, B), rng.uniform(10, 15, B), rng.uniform(0, 2 * np.pi, B)
    wind = rng.normal(0, 2.5, (B, 2))
    time = np.arange(0, T * 0.01 - 1e-9, 0.01)
    X0 = np.zeros((B, 5))
    for b in range(B):
        X0[b] = orc.flatness(orc.Circle([cx[b], cy[b]], r[b], v[b], alpha0=a0[b]).get(0.), wind[b])[0]
    X0 += rng.normal(0, 1, (B, 5)) * np.array([5, 5, 0.2, 0.05, 0.5])
    res = simulation.rollout(time, trajectory.CircleBatch(cx, cy, r, v, a0), wind, X0)
    assert not res.flags.any()
    worst = 0.
    for b in range(B):
        Xo, Uo, _, _, _ = orc.run_simulation(time, orc.Circle([cx[b], cy[b]], r[b], v[b], alpha0=a0[b]), wind[b], X0[b])
        worst = max(worst, np.abs(res.X[b] - Xo).max(), np.abs(res.U[b] - Uo).max())
    print("population worst |d| =", worst)
    assert worst < TOL


def test_formation_c2_against_reference_golden(d2d, golden):
    from d2d_b200 import simulation
    g = golden["formation"]
    X, U, time, _, _, Rr, eth = simulation.CircularFormationGVF(np.array([0, 0]), 60, 6, 60)
    assert len(time) == int(g["c2/T"])
    np.testing.assert_allclose(X[::4], g["c2/X"], rtol=0, atol=TOL)
    np.testing.assert_allclose(U[::4], g["c2/U"], rtol=0, atol=TOL)
    np.testing.assert_allclose(Rr[::4], g["c2/Rr"], rtol=0, atol=1e-8)
    np.testing.assert_allclose(eth[::4], g["c2/eth"], rtol=0, atol=1e-8)
    np.testing.assert_allclose(X[-1], g["c2/Xlast"], rtol=0, atol=TOL)


def test_formation_script09_and_reference_csv(d2d, golden):
    """Script 09's configuration (4 aircraft, own centres) and the reference's fixture src/states_over_time.csv."""
    from d2d_b200 import simulation
    g = golden["formation"]
    c9 = np.array([[0, -20], [25, -40], [25, -80], [0, -100]], dtype=float)
    X, U, time, _, _, Rr, eth = simulation.CircularFormationGVF(c9, 60, 4, 60, kd=25, z_des=np.zeros(3))
    np.testing.assert_allclose(X[::4], g["s09/X"], rtol=0, atol=TOL)
    np.testing.assert_allclose(Rr[::4], g["s09/Rr"], rtol=0, atol=1e-8)
    X, *_ = simulation.CircularFormationGVF(c9, 60, 4, 200, kd=25, z_des=np.zeros(3), nsub=1, tau_phi=0.9667)
    np.testing.assert_allclose(X[::20], g["csv_rk4_1/X"], rtol=0, atol=1e-8)
    assert np.abs(X[::20] - g["csv/X"]).max() < 2e-3         # vs the LSODA-integrated CSV: integrator gap only


def test_formation_batch_is_replicated_single(d2d):
    """Many formations in one launch (5 per warp for n_ac = 6) give the single-formation answer."""
    from d2d_b200 import simulation
    n_ac, F = 6, 37
    z = np.ones(n_ac - 1) * 2 * np.pi / n_ac
    X1 = np.array([20, 30, -np.pi / 2, 0, 10.])
    one = simulation.formation_rollout(np.zeros((1, n_ac, 2)), 60., n_ac, 200, 0.05, 4e-4, 15, 20, z, X1)
    many = simulation.formation_rollout(np.zeros((F, n_ac, 2)), 60., n_ac, 200, 0.05, 4e-4, 15, 20, z, X1)
    for f in range(F):
        np.testing.assert_array_equal(many["X"][f], one["X"][0])


def test_dcf_and_gvf_single_calls(d2d, golden):
    from d2d_b200 import guidance, simulation
    from oracle import d2d_oracle as orc
    g = golden["formation"]
    X = np.array([20, 30, -np.pi / 2, 0, 10.])
    tr = guidance.CircleTraj(np.array([0, -20]))
    e, n, H = tr.get(X, 60)
    U, U1, U2 = guidance.GVFcontroller(tr, None, None).get(X, 4e-4, 25, e, n, H)
    np.testing.assert_allclose([e, n[0], n[1], U, U1, U2], g["gvf_known"], rtol=1e-12)
    rng = np.random.default_rng(3)
    n_ac = 5
    p, c = rng.normal(0, 40, (2, n_ac)), rng.normal(0, 5, (n_ac, 2))
    B = simulation.chain_incidence(n_ac); z = np.ones(n_ac - 1) * 2 * np.pi / n_ac
    Ur, e_deg = guidance.DCFController().get(n_ac, B, c, p, z, 20)
    Uo, eo = orc.dcf(B, c, p, z, 20)
    np.testing.assert_allclose(Ur[:, 0], Uo, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(e_deg[:, 0], eo, rtol=1e-12, atol=1e-12)
    with pytest.raises(ValueError):
        guidance.DCFController().get(n_ac, B, np.array([0., 0.]), p, z, 20)      # 1-D centre, cf. SURVEY D2


def test_integration_md_ctypes_stub_runs(d2d, golden):
    """The reference-side ctypes stub printed in INTEGRATION.md, executed verbatim against libd2dx.so."""
    import os, re
    from d2d_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    md = open(os.path.join(root, "INTEGRATION.md")).read()
    code = [b for b in re.findall(r"```python\n(.*?)```", md, flags=re.S) if "run_simulation_gpu" in b][0]
    code = code.replace('C.CDLL("libd2dx.so")', f'C.CDLL({_lib.LIB_PATH!r})')
    ns = {}
    exec(code, ns)
    g = golden["dfff_c1"]
    X, U = ns["run_simulation_gpu"](g["time"], (30., 30., 30., 10., 3 * np.pi / 2, 0.), g["wind"], g["X0"])
    np.testing.assert_allclose(X, g["X"], rtol=0, atol=TOL)
    np.testing.assert_allclose(U, g["U"], rtol=0, atol=TOL)


def test_engine_elementary_functions_accuracy(d2d):
    """The straight-line sincos / atan2 / atan / div / sqrt / rsqrt of csrc/d2dx_math.cuh vs NumPy: <= 4 ulp."""
    eng = d2d.get_engine()
    rng = np.random.default_rng(9)
    n = 1 << 16
    x = np.concatenate([rng.uniform(-8, 8, n // 2), rng.uniform(-1e4, 1e4, n // 4), rng.normal(0, 1e-3, n // 8), rng.normal(0, 50, n // 8)])
    y = rng.normal(0, 5, n) * np.exp(rng.uniform(-6, 6, n))
    x[:8] = [0.4375, -0.4375, 0.6875, 1.1875, 2.4375, 1.0, -1.0, np.pi / 4]
    # the heading wrap's case boundaries: a = x + pi at 0, -0.0, 2 pi, 4 pi, the neighbouring doubles, a few laps out
    edge = np.array([-np.pi, np.pi, 3 * np.pi, -3 * np.pi, 5 * np.pi, -5 * np.pi, 0.0, -0.0, 1e-300, -1e-300, 40.0, -40.0, 1e6, -1e6])
    edge = np.concatenate([edge, np.nextafter(edge, np.inf), np.nextafter(edge, -np.inf), np.nextafter(np.nextafter(edge, -np.inf), -np.inf)])
    x[8:8 + len(edge)] = edge
    x[100:4196] = rng.uniform(-4 * np.pi, 4 * np.pi, 4096)
    out = eng.math_probe(eng.to_device(x), eng.to_device(y)).cpu().numpy()
    with np.errstate(all="ignore"):
        ref = [np.sin(x), np.cos(x), np.arctan2(y, x), np.arctan(x), y / x, np.sqrt(np.abs(x)), 1 / np.sqrt(np.abs(x)), 1 / x]
    names = ["sin", "cos", "atan2", "atan", "div", "sqrt", "rsqrt", "rcp"]
    wrapped = (x + np.pi) % (2 * np.pi) - np.pi                      # d2d/utils.py:7
    assert np.array_equal(out[10], wrapped), (x[out[10] != wrapped][:5], out[10][out[10] != wrapped][:5], wrapped[out[10] != wrapped][:5])
    # the two correction polynomials are truncated for seed residuals of 2^-20 (reciprocal) and 2^-19 (1 - a y^2): the error after
    # the correction is the cube of that
    seed = np.abs(out[8:10, 8 + len(edge):]).max(axis=1)
    print("seed residuals: rcp %.3g rsqrt %.3g (2^-20 = %.3g)" % (seed[0], seed[1], 2.**-20))
    assert seed[0] < 1.1 * 2.**-20 and seed[1] < 2.**-19, seed
    keep = np.ones(n, bool); keep[8:8 + len(edge)] = False          # the wrap's edge values are outside the other functions' domains
    x, y, out = x[keep], y[keep], out[:, keep]
    for k, (r, nm) in enumerate(zip(ref, names)):
        r = r[keep]
        ulp = np.abs(out[k] - r) / np.spacing(np.abs(r))
        if nm in ("sin", "cos"):            # absolute accuracy near zeros of sin/cos for large |x| is bounded by the 3-term reduction
            ok = (ulp <= 4) | (np.abs(out[k] - r) < 1e-15)
        else:
            ok = ulp <= (2 if nm in ('div', 'rcp', 'sqrt', 'rsqrt') else 4)
        assert ok.all(), (nm, float(ulp.max()), x[np.argmax(ulp)], y[np.argmax(ulp)])
        print(nm, "max ulp", float(ulp[ok].max()))


def test_minsnap_population_against_c_oracle(d2d):
    """C5's second population: randomised min-snap polynomials, POLY-specialised kernel, vs the C oracle."""
    from oracle import c_oracle as co, d2d_oracle as orc
    from d2d_b200 import simulation, trajectory
    rng = np.random.default_rng(99)
    B, dur = 300, 33.65
    Y0 = np.zeros((B, 2, 4)); Y1 = np.zeros((B, 2, 4))
    ang0, ang1 = rng.uniform(-0.5, 0.5, B), rng.uniform(1.0, 2.0, B)
    Y0[:, 0, 0], Y0[:, 1, 0] = rng.uniform(-20, 20, B), rng.uniform(-20, 20, B)
    Y0[:, 0, 1], Y0[:, 1, 1] = 10 * np.cos(ang0), 10 * np.sin(ang0)
    Y1[:, 0, 0], Y1[:, 1, 0] = Y0[:, 0, 0] + rng.uniform(150, 250, B), Y0[:, 1, 0] + rng.uniform(150, 250, B)
    Y1[:, 0, 1], Y1[:, 1, 1] = 10 * np.cos(ang1), 10 * np.sin(ang1)
    batch = trajectory.MinSnapBatch.from_boundaries(Y0, Y1, dur)
    # the vectorised coefficient solve agrees with the per-object one
    one = trajectory.MinSnapPoly(Y0[0], Y1[0], dur)
    np.testing.assert_allclose(batch.coefs0[0, 0], one._polys[0].coefs[0], rtol=1e-12, atol=1e-18)
    wind = rng.normal(0, 1.5, (B, 2))
    time = np.arange(0, 20., 0.01)
    par = np.zeros((B, 17)); par[:, 1:9] = batch.coefs0[:, 0]; par[:, 9:17] = batch.coefs0[:, 1]
    X0 = np.zeros((B, 5))
    for b in range(B):
        ms = orc.MinSnap(Y0[b], Y1[b], dur)
        for c in range(2):
            ms._polys[c].coefs[0] = batch.coefs0[b, c]
            for d in range(1, 4):
                for pw in range(8 - d):
                    ms._polys[c].coefs[d, pw] = orc._arr(d, pw + d) * batch.coefs0[b, c, pw + d]
        X0[b] = orc.flatness(ms.get(0.), wind[b])[0]
    X0 += rng.normal(0, 1, (B, 5)) * np.array([3, 3, 0.1, 0.02, 0.3])
    res = simulation.rollout(time, batch, wind, X0, log_every=10)
    ref = co.rollout(time, np.full(B, co.T_POLY, np.int32), par, wind, X0, log_every=10)
    assert ref["failed"] == 0 and not res.flags.any()
    ok = ref["max_err"] < 10.
    assert ok.sum() > 0.8 * B
    err = np.abs(res.X[ok] - ref["X"][ok]).max()
    print("min-snap population: max |dX| =", err, "over", int(ok.sum()), "scenarios")
    assert err < TOL


@pytest.mark.parametrize("tag", ["exp0", "exp13"])
def test_tabulated_trajectory_against_reference_golden(d2d, golden, tag, tmp_path):
    """TrajTabulated (planner solution -> tracker reference) on solutions shipped with the reference."""
    from d2d_b200 import simulation, trajectory_factory as ddtf
    g = golden["tabulated"]
    fn = tmp_path / "plan.npz"
    np.savez(fn, **{k: g[f"{tag}/{k}"] for k in ("sol_time", "sol_x", "sol_y", "sol_psi", "sol_phi", "sol_v", "wind")})
    traj = ddtf.TrajTabulated(str(fn))
    time = np.arange(0., traj.duration + 0.5, 0.01)
    assert len(time) == int(g[f"{tag}/T"])
    np.testing.assert_array_equal(traj.get_many(time[::7]), g[f"{tag}/Yref"])
    other = ddtf.TrajCircle()                                # a mixed batch: tabulated next to an analytic trajectory
    res = simulation.rollout(time, [other, traj], [0.5, -0.3], np.stack([np.array([60., 30., 1.5, 0., 10.]), g[f"{tag}/X0"]]))
    np.testing.assert_allclose(res.X[1][::5], g[f"{tag}/X"], rtol=0, atol=TOL)
    np.testing.assert_allclose(res.U[1][::5], g[f"{tag}/U"], rtol=0, atol=TOL)
    np.testing.assert_allclose(res.X_final[1], g[f"{tag}/Xlast"], rtol=0, atol=TOL)


def test_remaining_registry_scenarios(d2d, golden, tmp_path):
    """`oval` (d2d/scenario.py:268-279; upstream forgets Scenario.__init__, so the oracle is the checker) and the tabulated
    scenarios `dual opty` / `opty2` (:283-303) fed with planner files made from the golden solutions."""
    from oracle import d2d_oracle as orc
    from d2d_b200 import scenario, simulation
    scen, desc = scenario.get("oval")
    Xs, Us, _ = simulation.test_simulation(scen)
    Xo, Uo = orc.run_simulation(scen.time[:600], orc.traj_line_with_intro(Y0=(0., 100.), Y1=(0., 50.), Y2=(200., 50.), r=25.), [0., 2.5],
                                np.array(scen.X0s[0], float))[:2]
    np.testing.assert_allclose(Xs[0][:600], Xo, rtol=0, atol=TOL)
    np.testing.assert_allclose(Us[0][:599], Uo[:599], rtol=0, atol=TOL)
    g = golden["tabulated"]
    files = []
    for tag in ("exp0", "exp13", "exp0"):
        fn = tmp_path / f"plan_{len(files)}.npz"
        np.savez(fn, **{k: g[f"{tag}/{k}"] for k in ("sol_time", "sol_x", "sol_y", "sol_psi", "sol_phi", "sol_v", "wind")})
        files.append(str(fn))
    assert set(("oval", "dual opty", "opty2")) <= {e.split(":")[0] for e in scenario.list_available()}
    sc = scenario.ScenDualOpty(files)
    assert len(sc.trajs) == 3 and len(sc.X0s) == 3
    Xs, Us, Yrefs = simulation.test_simulation(sc)
    solo = simulation.rollout(sc.time, [sc.trajs[1]], sc.windfield.sample(0, None), np.asarray(sc.X0s[1], float)[None])
    np.testing.assert_allclose(Xs[1], solo.X[0], rtol=0, atol=1e-12)
    with pytest.raises(FileNotFoundError):                      # the upstream default files are not shipped
        scenario.get("opty2")


def test_step_by_step_duck_typed_api_under_d2d_alias(d2d, golden):
    """A caller written in the reference's style (`import d2d.dynamic as ddyn`, one controller call and one disc_dyn
    call per step, as the loop body of 05_test_simulation.py:28-32) runs against the package aliased as `d2d`."""
    import sys
    saved = {k: v for k, v in sys.modules.items() if k == "d2d" or k.startswith("d2d.")}
    try:
        for m in ("dynamic", "guidance", "trajectory", "trajectory_factory", "scenario", "utils", "opty_utils", "multiopty_utils"):
            sys.modules[f"d2d.{m}"] = __import__(f"d2d_b200.{m}", fromlist=[m])
        sys.modules["d2d"] = d2d
        import d2d.dynamic as ddyn
        import d2d.guidance as ddg
        import d2d.trajectory as ddt
        g = golden["dfff_c1"]
        time = g["time"][:60]
        traj, ac, wind = ddt.TrajectoryCircle(alpha0=3 * np.pi / 2), ddyn.Aircraft(), ddg.WindField([5, 0])
        ctl = ddg.DFFFController(traj, ac, wind)
        X, U = np.zeros((len(time), ddyn.Aircraft.s_size)), np.zeros((len(time), ddyn.Aircraft.i_size))
        X[0] = g["X0"]
        for i in range(1, len(time)):
            U[i - 1] = ctl.get(X[i - 1], time[i - 1])
            X[i] = ac.disc_dyn(X[i - 1], U[i - 1], wind, time[i - 1], time[i] - time[i - 1])
        np.testing.assert_allclose(X, g["X"][:60], rtol=0, atol=TOL)
        np.testing.assert_allclose(U[:-1], g["U"][:59], rtol=0, atol=TOL)
        np.testing.assert_allclose(np.array(ctl.K), g["K"][:59], rtol=0, atol=TOL)
        Yr = traj.get(0.5)
        assert Yr.shape == (4, 2)
        Xr, Ur, Xrd = ddg.DiffFlatness.state_and_input_from_output(Yr, wind.sample(0.5, Yr[0]), ac)
        assert Xr.shape == (5,) and Ur.shape == (2,)
    finally:
        for k in [k for k in sys.modules if k == "d2d" or k.startswith("d2d.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_mixed_population_grouped_by_family_equals_generic_kernel(d2d):
    """A mixed population is split into one launch per trajectory family (specialised kernels); the result equals the
    single generic-kernel launch in the caller's order."""
    from d2d_b200 import simulation, trajectory, trajectory_factory as ddtf
    rng = np.random.default_rng(21)
    trajs, X0 = [], []
    for k in range(420):
        kind = k % 4
        if kind == 0:
            tr = trajectory.TrajectoryCircle(c=[rng.uniform(-20, 20), rng.uniform(-20, 20)], r=rng.uniform(30, 60), v=rng.uniform(10, 12), alpha0=rng.uniform(0, 6))
        elif kind == 1:
            tr = trajectory.MinSnapPoly([[0, 10, 0, 0], [rng.uniform(-5, 5), 0, 0, 0]], [[200, 0, 0, 0], [200, 10, 0, 0]], duration=33.65)
        elif kind == 2:
            tr = trajectory.TrajectoryLine([0, rng.uniform(0, 30)], [100, 25], v=10.)
        else:
            tr = ddtf.TrajSquare()
        trajs.append(tr)
        X0.append([rng.uniform(-2, 2), rng.uniform(-2, 2), rng.uniform(-0.3, 0.3), 0., 10.])
    X0 = np.array(X0)
    for b, tr in enumerate(trajs):
        if isinstance(tr, trajectory.TrajectoryCircle):
            X0[b, :2] += tr.c + np.array([tr.r * np.cos(tr.alpha0), tr.r * np.sin(tr.alpha0)]); X0[b, 2] += tr.alpha0 + np.pi / 2
    time = np.arange(0, 3., 0.01)
    wind = rng.normal(0, 1., (len(trajs), 2))
    a = simulation.rollout(time, trajs, wind, X0, log_every=5, log_ref=True)
    b = simulation.rollout(time, trajs, wind, X0, log_every=5, log_ref=True, group_by_family=False)
    for name in ("X", "U", "Xref", "K", "X_final"):
        np.testing.assert_allclose(getattr(a, name), getattr(b, name), rtol=0, atol=1e-12, err_msg=name)
    for name in ("sum_sq_err", "max_err"):
        np.testing.assert_allclose(getattr(a, name), getattr(b, name), rtol=1e-12, err_msg=name)
    np.testing.assert_allclose(a.pop_sum_sq_err, b.pop_sum_sq_err, rtol=1e-12)
    assert a.pop_max_err == b.pop_max_err and not a.flags.any()


@pytest.mark.parametrize("nsub,tau_phi,dt", [(3, 0.01, 0.01), (1, 0.9667, 0.02), (10, 0.01, 0.1)])
def test_dfff_rollout_substeps_and_time_constants_against_oracle(d2d, nsub, tau_phi, dt):
    """RK4 sub-stepping, other time constants and coarser control periods in the DFFF rollout (the fast stage-heading
    path must hand over to the generic one when the increments grow)."""
    from oracle import d2d_oracle as orc
    from d2d_b200 import simulation, trajectory
    time = np.arange(0, 150 * dt - 1e-9, dt)
    specs = [dict(c=[30., 30.], r=30., v=10., alpha0=4.0), dict(c=[0., 0.], r=-25., v=12., alpha0=1.0)]
    X0 = np.array([[12., 5., 1.0, 0.1, 9.], [-20., 18., -2.0, 0., 12.]])
    wind = [1.5, -1.0]
    res = simulation.rollout(time, [trajectory.TrajectoryCircle(**s) for s in specs], wind, X0, tau_phi=tau_phi, tau_v=1.3, nsub=nsub)
    assert not res.flags.any()
    for b, s in enumerate(specs):
        Xo, Uo, _, _, _ = orc.run_simulation(time, orc.Circle(**s), wind, X0[b], nsub=nsub, tau_phi=tau_phi, tau_v=1.3)
        np.testing.assert_allclose(res.X[b], Xo, rtol=0, atol=TOL)
        np.testing.assert_allclose(res.U[b], Uo, rtol=0, atol=TOL)


def test_small_helpers_run_on_the_engine(d2d, golden):
    """norm_mpi_pi and CircleTraj.get go through the C ABI as well (no host arithmetic on the path)."""
    from d2d_b200 import guidance, utils
    u = golden["units"]
    np.testing.assert_array_equal(guidance.norm_mpi_pi(u["ang"]), u["wrapped"])          # bit-identical to NumPy's floored %
    assert guidance.norm_mpi_pi(4.0) == (4.0 + np.pi) % (2 * np.pi) - np.pi and utils.norm_mpi_pi is guidance.norm_mpi_pi
    e, n, H = guidance.CircleTraj(np.array([0, -20])).get(np.array([20, 30, -np.pi / 2, 0, 10.]), 60)
    assert float(e) == -700.0 and list(n) == [40.0, 100.0] and H.tolist() == [[2, 0], [0, 2]]
    l0 = d2d.get_engine().launches
    guidance.norm_mpi_pi(np.zeros(3)); guidance.CircleTraj().get(np.zeros(5), 1.)
    assert d2d.get_engine().launches == l0 + 2


def _ring_incidence(n):
    B = np.zeros((n, n))
    for j in range(n):
        B[j, j], B[(j + 1) % n, j] = -1, 1
    return B


def _complete_incidence(n):
    edges = [(a, b) for a in range(n) for b in range(a + 1, n)]
    B = np.zeros((n, len(edges)))
    for k, (a, b) in enumerate(edges):
        B[a, k], B[b, k] = -1, 1
    return B


@pytest.mark.parametrize("graph,n_ac", [("ring", 5), ("complete", 4), ("complete", 6)])
def test_dcf_general_incidence_matrices(d2d, graph, n_ac):
    """The reference's DCFController.get takes ANY incidence matrix (d2d/guidance.py:103-126): ring graphs (n_e = n_ac) and
    complete graphs (n_e > n_ac: lanes own several edges) in the single call and in the formation rollout, several
    formations per warp."""
    from d2d_b200 import guidance, simulation
    from oracle import d2d_oracle as orc
    rng = np.random.default_rng(30 + n_ac)
    B = _ring_incidence(n_ac) if graph == "ring" else _complete_incidence(n_ac)
    n_e = B.shape[1]
    z = rng.uniform(-1, 1, n_e)
    p, c = rng.normal(0, 40, (2, n_ac)), rng.normal(0, 5, (n_ac, 2))
    Ur, e_deg = guidance.DCFController().get(n_ac, B, c, p, z, 20)
    Uo, eo = orc.dcf(B, c, p, z.copy(), 20)
    np.testing.assert_allclose(Ur[:, 0], Uo, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(np.ravel(e_deg), eo, rtol=1e-12, atol=1e-12)
    F, T = 7, 120
    cs = rng.normal(0, 3, (F, n_ac, 2))
    X1 = np.array([20, 30, -np.pi / 2, 0, 10.])
    out = simulation.formation_rollout(cs, 60., n_ac, T, 0.05, 4e-4, 15, 5., z, X1, B=B)
    for f in (0, F - 1):
        Xo, Uo_, _, Rro, etho = orc.run_formation(cs[f], 60., n_ac, T * 0.05 - 1e-9, 4e-4, 15, 5., z.copy(), B=B)
        np.testing.assert_allclose(out["X"][f], Xo, rtol=0, atol=1e-9)
        np.testing.assert_allclose(out["e_theta"][f], etho, rtol=0, atol=1e-8)


def test_rollout_without_final_control_has_defined_last_row(d2d, golden):
    """final_control=False skips the trailing ctl.get of 05_test_simulation.py:33: the last state row is still the state at
    the last sample, the last input row reads zero (never uninitialised memory)."""
    from d2d_b200 import simulation, trajectory
    g = golden["dfff_c1"]
    time = g["time"][:300]
    tr = trajectory.TrajectoryCircle(alpha0=3 * np.pi / 2)
    a = simulation.rollout(time, [tr], [5., 0.], g["X"][:1], log_ref=True)
    b = simulation.rollout(time, [tr], [5., 0.], g["X"][:1], log_ref=True, final_control=False)
    np.testing.assert_array_equal(b.X, a.X)
    np.testing.assert_array_equal(b.U[:, :-1], a.U[:, :-1])
    assert not b.U[:, -1].any() and not b.K[:, -1].any() and not b.Xref[:, -1].any()
    np.testing.assert_array_equal(b.X[:, -1], b.X_final)
