"""The oracle against the golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import d2d_oracle as orc

TOL = 1e-11     # oracle (NumPy restatement) vs reference output; both run the same libm / SciPy


def test_units_flatness_jac_lqr(golden):
    u = golden["units"]
    for i in range(len(u["Ys"])):
        Xr, Ur, Xd = orc.flatness(u["Ys"][i], u["Ws"][i])
        np.testing.assert_allclose(Xr, u["Xr"][i], rtol=0, atol=1e-13)
        np.testing.assert_allclose(Ur, u["Ur"][i], rtol=0, atol=1e-13)
        np.testing.assert_allclose(Xd, u["Xrdot"][i], rtol=0, atol=1e-13)
        A1, B1 = orc.cont_jac_3(Xr)
        np.testing.assert_array_equal(A1, u["A"][i][:3, :3])
        np.testing.assert_array_equal(B1, u["A"][i][:3, 3:])
        np.testing.assert_allclose(orc.lqr(A1, B1, orc._Q, orc._R), u["K"][i], rtol=1e-12, atol=1e-13)
        np.testing.assert_array_equal(orc.cont_dyn(u["Xs"][i], u["Us"][i], u["Ws"][i]), u["Xdot"][i])
    np.testing.assert_array_equal(orc.norm_mpi_pi(u["ang"]), u["wrapped"])


def test_minsnap_coefficients(golden):
    u = golden["units"]
    ms = orc.traj_minsnap_demo()
    np.testing.assert_array_equal(np.array([p.coefs for p in ms._polys]), u["minsnap_coefs"])
    np.testing.assert_array_equal(ms.get(10.0), u["minsnap_get10"])
    # SURVEY appendix C known answer
    assert abs(ms.get(10.0)[0, 0] - 99.43317456065962) < 1e-12


def test_c1_closed_loop(golden):
    g = golden["dfff_c1"]
    X, U, Yref, Xref, K = orc.run_simulation(g["time"], orc.Circle(alpha0=3 * np.pi / 2), g["wind"], g["X0"])
    np.testing.assert_allclose(X, g["X"], rtol=0, atol=TOL)
    np.testing.assert_allclose(U, g["U"], rtol=0, atol=TOL)
    np.testing.assert_allclose(Xref, g["Xref"], rtol=0, atol=TOL)
    np.testing.assert_allclose(K, g["K"][:, :, :3], rtol=0, atol=TOL)
    np.testing.assert_array_equal(Yref, g["Yref"])
    # SURVEY appendix C known answers
    np.testing.assert_allclose(X[999], [24.411920311867036, 59.467375958777055, -3.0166728059917847,
                                        0.32821028964231175, 14.974098817914387], atol=1e-10)
    np.testing.assert_allclose(U[0], [0.7853981633974483, 4.0], atol=0)


SCENS = ["line", "line2", "square", "mucir", "mucir2", "patrol", "patrol_2", "patrol_3", "circForm"]


@pytest.mark.parametrize("name", SCENS)
def test_scenario_registry(golden, name):
    g = golden["dfff_scenarios"]
    s = orc.scenario(name)
    # trajectories (Yref) for every aircraft, closed loop for the first and last only (CPU time)
    n = len(s["trajs"])
    assert int(g[f"{name}/0/T"]) == len(s["time"])
    for i in range(n):
        Y = np.array([s["trajs"][i].get(t) for t in s["time"][::25]])
        np.testing.assert_allclose(Y, g[f"{name}/{i}/Yref"], rtol=0, atol=1e-12)
    for i in sorted({0, n - 1}):
        T = min(len(s["time"]), 1001)          # the first 1000 steps pin the loop; full length is in the C oracle test
        X, U, _, _, K = orc.run_simulation(s["time"][:T], s["trajs"][i], s["wind"], s["X0s"][i], s["perts"][i][:T])
        np.testing.assert_allclose(X[::5], g[f"{name}/{i}/X"][:len(X[::5])], rtol=0, atol=1e-10)
        np.testing.assert_allclose(U[:-1:5], g[f"{name}/{i}/U"][:len(U[:-1:5])], rtol=0, atol=1e-10)


@pytest.mark.parametrize("name", ["minsnap", "sidemo", "slalom"])
def test_extra_trajectories(golden, name):
    g = golden["dfff_scenarios"]
    traj = {"minsnap": orc.traj_minsnap_demo, "sidemo": orc.traj_si_demo, "slalom": orc.Slalom}[name]()
    time = np.arange(0., traj.duration, 0.01)
    assert len(time) == int(g[f"{name}/0/T"])
    Y = np.array([traj.get(t) for t in time[::25]])
    np.testing.assert_allclose(Y, g[f"{name}/0/Yref"], rtol=0, atol=1e-12)
    T = 501
    X, U, _, _, _ = orc.run_simulation(time[:T], traj, g[f"{name}/0/wind"], g[f"{name}/0/X0"])
    np.testing.assert_allclose(X[::5], g[f"{name}/0/X"][:len(X[::5])], rtol=0, atol=1e-10)


def test_formation_c2(golden):
    g = golden["formation"]
    n_ac = 6
    X, U, time, Rr, eth = orc.run_formation(np.zeros((n_ac, 2)), 60, n_ac, 60, 4e-4, 15, 20,
                                            np.ones(n_ac - 1) * 2 * np.pi / n_ac, nsub=5)
    assert len(time) == int(g["c2/T"])
    np.testing.assert_allclose(X[::4], g["c2/X"], rtol=0, atol=1e-10)
    np.testing.assert_allclose(U[::4], g["c2/U"], rtol=0, atol=1e-10)
    np.testing.assert_allclose(Rr[::4], g["c2/Rr"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(eth[::4], g["c2/eth"], rtol=0, atol=1e-9)
    # SURVEY appendix C
    np.testing.assert_allclose(X[1, 0], [20.009981682455503, 29.494008318772817, -1.5224454584742235,
                                         0.9269760751347237, 10.24385287747645], atol=1e-12)


def test_formation_script09_and_csv(golden):
    """Script 09 (4 aircraft, own centres) under RK4, and the reference's own fixture
    src/states_over_time.csv (LSODA, tau_phi = 0.9667): RK4 agrees with it to LSODA's tolerance class."""
    g = golden["formation"]
    c9 = np.array([[0, -20], [25, -40], [25, -80], [0, -100]], dtype=float)
    X, U, time, Rr, eth = orc.run_formation(c9, 60, 4, 60, 4e-4, 25, 20, np.zeros(3), nsub=5)
    np.testing.assert_allclose(X[::4], g["s09/X"], rtol=0, atol=1e-10)
    X, _, _, _, _ = orc.run_formation(c9, 60, 4, 200, 4e-4, 25, 20, np.zeros(3), nsub=1, tau_phi=0.9667)
    np.testing.assert_allclose(X[::20], g["csv_rk4_1/X"], rtol=0, atol=1e-9)
    assert np.abs(X[::20] - g["csv/X"]).max() < 2e-3        # vs the CSV itself: integrator gap (7.7e-4 measured)
    assert np.abs(X[1] - g["csv/row1"]).max() < 1e-6
    e, n0, n1, Ug, U1g, U2g = g["gvf_known"]
    Uo, U1o, U2o = orc.gvf(np.array([20, 30, -np.pi / 2, 0, 10.]), [0, -20], 60, 4e-4, 25)
    np.testing.assert_allclose([Uo, U1o, U2o], [Ug, U1g, U2g], rtol=1e-13)
    assert abs(Ug - 25.15764653581325) < 1e-12


def test_planner_timing_triangle(golden):
    g = golden["colloc"]
    args = [(0, 20, 50), (0, 19.98, 50), (0, 9.98, 50), (0, 4.2, 50), (0, 5.5, 10), (0, 10, 10), (0, 0.29, 100), (2., 17., 50.)]
    for a, ref in zip(args, g["planner_timing"]):
        np.testing.assert_allclose(orc.planner_timing(*a), ref, rtol=0, atol=0)
    np.testing.assert_array_equal(np.array(orc.triangle([0., 0.], [50., 0.], 12., 4.2, 7, go_left=-1)), g["triangle7"])


def _inst(g, tag):
    return [(int(k), int(n), v) for (k, n, v) in g[f"{tag}/inst"]]


@pytest.mark.parametrize("tag", ["c3", "c3w"])
def test_colloc_single(golden, tag):
    g = golden["colloc"]
    N, h = 1001, 0.02
    free, inst = g["c3/free"], _inst(g, "c3")
    res = orc.colloc_residual(free, N, 1, h, g[f"{tag}/wind"], inst)
    np.testing.assert_allclose(res, g[f"{tag}/residual"], rtol=1e-12, atol=1e-12)
    Jd = orc.colloc_jac_dense(free, N, 1, h)
    np.testing.assert_allclose(Jd, g[f"{tag}/jac_dense"].reshape(-1), rtol=1e-12, atol=0)
    rows, cols = orc.colloc_structure(N, 1, [], "dense")
    np.testing.assert_array_equal(rows, g["c3/rows"].reshape(-1))
    np.testing.assert_array_equal(cols, g["c3/cols"].reshape(-1))
    # compact layout scattered through its structure equals the dense one
    rc, cc = orc.colloc_structure(N, 1, [], "compact")
    D = np.zeros((3 * (N - 1), 5 * N)); D[rc, cc] = orc.colloc_jac_compact(free, N, 1, h).reshape(-1)
    D2 = np.zeros_like(D); np.add.at(D2, (rows, cols), Jd)
    np.testing.assert_array_equal(D, D2)
    # the cached IPOPT solution satisfies the backward-Euler defects (pins the residual convention)
    r0 = orc.colloc_residual(g["c3/sol"], N, 1, h, [0., 0.], inst)
    assert np.abs(r0).max() < 1e-6


@pytest.mark.parametrize("tag,n_ac,N,h", [("m3", 3, 20, 0.1), ("c4", 16, 500, 0.02)])
def test_colloc_multi(golden, tag, n_ac, N, h):
    g = golden["colloc"]
    free, inst = g[f"{tag}/free"], _inst(g, tag)
    res = orc.colloc_residual(free, N, n_ac, h, g[f"{tag}/wind"], inst)
    np.testing.assert_allclose(res, g[f"{tag}/residual"], rtol=1e-12, atol=1e-11)
    Jd = orc.colloc_jac_dense(free, N, n_ac, h).reshape(N - 1, 3 * n_ac, 8 * n_ac)
    nz = Jd[:, g[f"{tag}/nz_eq"], g[f"{tag}/nz_col"]]
    np.testing.assert_allclose(nz, g[f"{tag}/jac_nz"], rtol=1e-12, atol=0)
    assert np.count_nonzero(Jd) <= 12 * n_ac * (N - 1)
    if tag == "m3":
        np.testing.assert_allclose(Jd, g["m3/jac_dense"], rtol=1e-12, atol=0)


SINGLE_SPECS = {
    "airvel": dict(vsp=12., kvel=1.),
    "bank": dict(kbank=1.),
    "input": dict(vsp=12., kvel=1., kbank=50.),
    "obs0": dict(kobs=1., obstacles=[(30, 0, 15.)], obs_kind=0),
    "obs1": dict(kobs=1., obstacles=[(5, 15, 10.)], obs_kind=1),
    "composit": dict(vsp=15., kvel=.5, kbank=1., kobs=.5, obstacles=[(5, 15, 10)], obs_kind=0),
    "composit1": dict(vsp=12., kvel=.7, kbank=1.5, kobs=2., obstacles=[(5, 15, 10), (-3., 20., 6.)], obs_kind=1),
}


@pytest.mark.parametrize("name", sorted(SINGLE_SPECS))
def test_costs_single(golden, name):
    g = golden["colloc"]
    for tag, free in (("sol", g["c3/sol"]), ("noisy", g["c3/free"])):
        c, gr = orc.cost_and_grad(free, 1001, 1, SINGLE_SPECS[name])
        np.testing.assert_allclose(c, g[f"cost1/{name}/{tag}/cost"], rtol=1e-12)
        np.testing.assert_allclose(gr, g[f"cost1/{name}/{tag}/grad"], rtol=1e-12, atol=1e-300)
    if name == "input":
        c, gr = orc.cost_and_grad(g["c3/free"], 1001, 1, dict(SINGLE_SPECS[name], obj_scale=3.5))
        np.testing.assert_allclose(c, g["cost1/input_scaled/noisy/cost"], rtol=1e-12)
        np.testing.assert_allclose(gr, g["cost1/input_scaled/noisy/grad"], rtol=1e-12)


MULTI_SPECS = {
    "input": dict(vsp=12., kvel=70., kbank=1.),
    "airvel": dict(vsp=12., kvel=1.),
    "bank": dict(kbank=1.),
    "obs0": dict(kobs=1., obstacles=[(60., 5., 12.)], obs_kind=0),
    "obs1": dict(kobs=1., obstacles=[(60., 5., 12.)], obs_kind=1),
    "collision": dict(kcol=1., rcol=10.),
    "composit": dict(vsp=12., kvel=70., kbank=1., kobs=0.5, kcol=10., obstacles=[(60., 5., 12.), (-20., 30., 8.)],
                     obs_kind=1, rcol=10.),
    "composit_nocol": dict(vsp=12., kvel=2., kbank=1.),
}


@pytest.mark.parametrize("tag,n_ac,N", [("m3", 3, 20), ("c4", 16, 500)])
@pytest.mark.parametrize("name", sorted(MULTI_SPECS))
def test_costs_multi(golden, tag, n_ac, N, name):
    g = golden["colloc"]
    free = g[f"{tag}/free"]
    c, gr = orc.cost_and_grad(free, N, n_ac, MULTI_SPECS[name], multi=True)
    np.testing.assert_allclose(c, g[f"{tag}/cost/{name}/cost"], rtol=1e-12)
    ref = g[f"{tag}/cost/{name}/grad"]
    if len(ref) != len(gr):
        gr = gr[::7]
    np.testing.assert_allclose(gr, ref, rtol=1e-12, atol=1e-300)


def test_collision_all_pairs_gradient_is_exact_derivative():
    """in exact_grad mode the all-pairs gradient is the true derivative of the cost."""
    rng = np.random.default_rng(7)
    N, n_ac = 6, 4
    free = rng.normal(0, 3., 5 * n_ac * N)
    spec = dict(kcol=10., rcol=10., pairs="all", exact_grad=True)
    c0, gr = orc.cost_and_grad(free, N, n_ac, spec, multi=True)
    for idx in rng.choice(3 * n_ac * N, 12, replace=False):
        d = np.zeros_like(free); d[idx] = 1e-6
        fd = (orc.cost_and_grad(free + d, N, n_ac, spec, multi=True)[0] - orc.cost_and_grad(free - d, N, n_ac, spec, multi=True)[0]) / 2e-6
        assert abs(fd - gr[idx]) < 1e-7 * max(1., abs(fd))


def test_sorted_inputs_12(golden):
    names = [str(s) for s in golden["colloc"]["sorted_inputs_12"]]
    assert names[:4] == ["phi0(t)", "phi1(t)", "phi10(t)", "phi11(t)"]


@pytest.mark.parametrize("tag", ["exp0", "exp13"])
def test_tabulated_trajectory(golden, tag):
    g = golden["tabulated"]
    tr = orc.Tabulated(g[f"{tag}/sol_time"], g[f"{tag}/sol_x"], g[f"{tag}/sol_y"], g[f"{tag}/sol_psi"], g[f"{tag}/sol_v"], g[f"{tag}/wind"])
    T = int(g[f"{tag}/T"])
    time = np.arange(0., tr.duration + 0.5, 0.01)
    assert len(time) == T
    np.testing.assert_array_equal(np.array([tr.get(t) for t in time[::7]]), g[f"{tag}/Yref"])
    n = 401
    X, U, _, _, _ = orc.run_simulation(time[:n], tr, [0.5, -0.3], g[f"{tag}/X0"])
    np.testing.assert_allclose(X[::5], g[f"{tag}/X"][:len(X[::5])], rtol=0, atol=1e-10)


@pytest.mark.parametrize("tag", ["hf", "stline", "inf"])
def test_tracker_other_outputs(golden, tag):
    g = golden["tracker"]
    X, U, *_ = orc.run_tracker(g[f"{tag}/time"], g[f"{tag}/x_ref"], g[f"{tag}/y_ref"], g[f"{tag}/wind"], g[f"{tag}/X0s"])
    np.testing.assert_allclose(X, g[f"{tag}/X"], rtol=0, atol=1e-10)
    np.testing.assert_allclose(U, g[f"{tag}/U"], rtol=0, atol=1e-10)


@pytest.mark.parametrize("tag", ["simple", "opt"])
def test_tracker_5state_lqr(golden, tag):
    """Controllers.DiffController + implement_controller vs the unmodified reference."""
    g = golden["tracker"]
    X, U, Xr, dX, K = orc.run_tracker(g[f"{tag}/time"], g[f"{tag}/x_ref"], g[f"{tag}/y_ref"], g[f"{tag}/wind"], g[f"{tag}/X0s"])
    np.testing.assert_allclose(X, g[f"{tag}/X"], rtol=0, atol=1e-10)
    np.testing.assert_allclose(U, g[f"{tag}/U"], rtol=0, atol=1e-10)
    np.testing.assert_allclose(Xr, g[f"{tag}/Xr"], rtol=0, atol=1e-11)
    np.testing.assert_allclose(dX, g[f"{tag}/dX"], rtol=0, atol=1e-10)
    np.testing.assert_allclose(K, g[f"{tag}/K"], rtol=1e-9, atol=1e-10)


def test_flatness5_single_calls(golden):
    g = golden["tracker"]
    for i in range(len(g["flat/Y"])):
        Y = g["flat/Y"][i]
        Xr, Ur = orc.flatness5(Y[0], Y[1], Y[2], Y[3], g["flat/W"][i])
        np.testing.assert_allclose(Xr, g["flat/Xr"][i], rtol=1e-13, atol=1e-13)
        np.testing.assert_allclose(Ur, g["flat/Ur"][i], rtol=1e-12, atol=1e-13)
