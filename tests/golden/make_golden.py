#!/usr/bin/env python3
"""Generate the committed golden vectors by running the UNMODIFIED reference.

Runs only where the reference checkout is mounted (default /root/reference,
override with $D2D_REF).  It imports the reference's own `d2d` package and
numbered scripts (with import-only stand-ins for matplotlib / control / opty
from `_refstubs/`), drives them on the configurations of SURVEY.md section 8(d),
and writes small `.npz` fixtures next to this file.  Nothing under `tests/`,
`bench.py` or the product package reads the reference at run time: they read
these fixtures.

What is reference code and what is ours here:
  * reference, unmodified: Aircraft.cont_dyn / cont_jac, DFFFController,
    DiffFlatness, every Trajectory class, every Scenario, run_simulation
    (05_test_simulation.py:21-34), CircularFormationGVF of script 09,
    DCFController / CircleTraj / GVFcontroller, the sympy EoM (get_eom),
    every cost class, planner_timing, triangle.
  * ours: the fixed-step RK4/ZOH replacement of Aircraft.disc_dyn (the
    reference integrates with adaptive LSODA; north_star defines parity
    against "the same fixed-step integrator and dt"), the C2 formation loop
    with a per-aircraft centre array (script 08 crashes for n_ac != 2), and the
    backward-Euler discretisation of the reference EoM (opty is not installed).
"""
import importlib.util
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("D2D_REF", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "_refstubs"))
sys.path.insert(0, os.path.join(REF, "src"))
warnings.filterwarnings("ignore", category=SyntaxWarning)

import scipy.integrate  # noqa: E402
import sympy as sym  # noqa: E402

import d2d.dynamic as ddyn  # noqa: E402
import d2d.guidance as ddg  # noqa: E402
import d2d.multiopty_utils as d2mou  # noqa: E402
import d2d.opty_utils as d2ou  # noqa: E402
import d2d.scenario as dds  # noqa: E402
import d2d.trajectory as ddt  # noqa: E402
import d2d.trajectory_factory as ddtf  # noqa: E402
import d2d.utils as d2u  # noqa: E402


def load_script(fname, modname):
    spec = importlib.util.spec_from_file_location(modname, os.path.join(REF, "src", fname))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


# --------------------------------------------------------------------------
# fixed-step RK4 / zero-order-hold replacement of Aircraft.disc_dyn
# --------------------------------------------------------------------------
_lsoda_disc_dyn = ddyn.Aircraft.disc_dyn


def make_rk4_disc_dyn(nsub):
    def disc_dyn(self, Xk, Uk, W, t, dt):
        X = np.array(Xk, dtype=float)
        h = dt / nsub
        for s in range(nsub):
            ts = t + s * h
            k1 = np.array(self.cont_dyn(X, ts, Uk, W))
            k2 = np.array(self.cont_dyn(X + 0.5 * h * k1, ts + 0.5 * h, Uk, W))
            k3 = np.array(self.cont_dyn(X + 0.5 * h * k2, ts + 0.5 * h, Uk, W))
            k4 = np.array(self.cont_dyn(X + h * k3, ts + h, Uk, W))
            X = X + (h / 6.0) * (k1 + 2.0 * k2 + 2.0 * k3 + k4)
        X[self.s_psi] = d2u.norm_mpi_pi(X[self.s_psi])
        return X
    return disc_dyn


def use_rk4(nsub):
    ddyn.Aircraft.disc_dyn = make_rk4_disc_dyn(nsub)


def use_lsoda():
    ddyn.Aircraft.disc_dyn = _lsoda_disc_dyn


# --------------------------------------------------------------------------
# path A: DFFF closed loop
# --------------------------------------------------------------------------
def run_dfff(sim, time, traj, wind, X0, perts, ac=None):
    ac = ac or ddyn.Aircraft()
    ctl = ddg.DFFFController(traj, ac, wind)
    X, U, Yref = sim.run_simulation(time, ac, wind, ctl, np.array(X0, dtype=float), perts)
    Xref = np.array(ctl.Xref)
    K = np.array(ctl.K)
    return X, U, Yref, Xref, K


def golden_c1(sim):
    traj = ddt.TrajectoryCircle(alpha0=3 * np.pi / 2)   # = ScenCircle(cst_gvel=True), scenario.py:106-107
    wind = ddg.WindField([5, 0])                          # scenario.py:114
    time = np.arange(0, 10, 0.01)
    ac = ddyn.Aircraft()
    Xr0 = ddg.DiffFlatness.state_and_input_from_output(traj.get(0.), wind.sample(0, None), ac)[0]
    X0 = Xr0 + np.array([5., -5., 0., 0., 0.])
    perts = np.zeros((len(time), 5))
    use_rk4(1)
    X, U, Yref, Xref, K = run_dfff(sim, time, traj, wind, X0, perts)
    use_lsoda()
    Xl, Ul, _, _, _ = run_dfff(sim, time, traj, wind, X0, perts)
    use_rk4(1)
    gap = np.abs(Xl - X).max(axis=0)
    print("C1: X[999] =", X[999], " LSODA-vs-RK4 max gap per state", gap)
    np.savez_compressed(os.path.join(HERE, "dfff_c1.npz"), time=time, X0=X0, X=X, U=U, Yref=Yref,
                        Xref=Xref[:len(time)], K=K[:len(time)], wind=np.array([5., 0.]),
                        lsoda_gap=gap, X_lsoda_final=Xl[-1])


def golden_scenarios(sim):
    """Every runnable scenario of the registry under DFFF + RK4(h=dt)."""
    use_rk4(1)
    out = {}
    names = ["line", "line2", "square", "mucir", "mucir2", "patrol", "patrol_2", "patrol_3", "circForm"]
    for name in names:
        scen, _ = dds.get(name)
        for i, (traj, X0, pert) in enumerate(zip(scen.trajs, scen.X0s, scen.perts)):
            X, U, Yref, Xref, K = run_dfff(sim, scen.time, traj, scen.windfield, X0, pert)
            key = f"{name}/{i}"
            out[key + "/X"] = X[::5]
            out[key + "/U"] = U[::5]
            out[key + "/Xlast"] = X[-1]
            out[key + "/Ulast"] = U[-1]
            out[key + "/Yref"] = Yref[::25]
            out[key + "/K"] = K[:len(scen.time):25]
            out[key + "/T"] = np.array(len(scen.time))
            assert np.isfinite(X).all(), key
            print(f"scenario {key}: T={len(scen.time)} X[-1]={X[-1]}")
    # extra trajectories not reachable through the scenario registry
    extras = {
        "minsnap": (ddtf.TrajMinSnapDemo(), [0., 0.], None),
        "sidemo": (ddtf.TrajSiDemo(), [0., 1.], None),
        "slalom": (ddtf.TrajSlalom(), [1., -1.], None),
    }
    for name, (traj, w, _) in extras.items():
        wind = ddg.WindField(w)
        time = np.arange(0., traj.duration, 0.01)
        ac = ddyn.Aircraft()
        X0 = ddg.DiffFlatness.state_and_input_from_output(traj.get(0.), w, ac)[0] + np.array([2., -3., 0.1, 0., 0.5])
        perts = np.zeros((len(time), 5))
        X, U, Yref, Xref, K = run_dfff(sim, time, traj, wind, X0, perts)
        key = f"{name}/0"
        out[key + "/X"] = X[::5]; out[key + "/U"] = U[::5]
        out[key + "/Xlast"] = X[-1]; out[key + "/Ulast"] = U[-1]
        out[key + "/Yref"] = Yref[::25]; out[key + "/K"] = K[:len(time):25]
        out[key + "/T"] = np.array(len(time)); out[key + "/X0"] = X0; out[key + "/wind"] = np.array(w)
        assert np.isfinite(X).all(), key
        print(f"extra {key}: T={len(time)} X[-1]={X[-1]}")
    np.savez_compressed(os.path.join(HERE, "dfff_scenarios.npz"), **out)


def golden_units():
    """Single-call known answers: flatness, cont_dyn, cont_jac, LQR gain, min-snap coefficients."""
    rng = np.random.default_rng(2024)
    ac = ddyn.Aircraft()
    n = 64
    Ys = rng.normal(0., 1., (n, 4, 2)) * np.array([50., 8., 2., 0.5])[None, :, None]
    Ws = rng.normal(0., 2., (n, 2))
    Xr = np.zeros((n, 5)); Ur = np.zeros((n, 2)); Xd = np.zeros((n, 5))
    A = np.zeros((n, 5, 5)); K = np.zeros((n, 2, 3))
    import control
    for i in range(n):
        Xr[i], Ur[i], Xd[i] = ddg.DiffFlatness.state_and_input_from_output(Ys[i], Ws[i], ac)
        A[i], _ = ac.cont_jac(Xr[i], Ur[i], 0., None)
        K[i] = control.lqr(A[i][:3, :3], A[i][:3, 3:], np.diag([1, 1, 0.1]), np.diag([8, 1]))[0]
    Xs = rng.normal(0., 1., (n, 5)) * np.array([50., 50., 2., 0.4, 1.]) + np.array([0, 0, 0, 0, 12.])
    Us = rng.normal(0., 1., (n, 2)) * np.array([0.3, 2.]) + np.array([0., 12.])
    Xdot = np.array([ac.cont_dyn(Xs[i], 0., Us[i], ddg.WindField(list(Ws[i]))) for i in range(n)])
    ang = rng.uniform(-30, 30, 256)
    wrapped = d2u.norm_mpi_pi(ang)
    ms = ddtf.TrajMinSnapDemo()
    np.savez_compressed(os.path.join(HERE, "units.npz"), Ys=Ys, Ws=Ws, Xr=Xr, Ur=Ur, Xrdot=Xd, A=A, K=K,
                        Xs=Xs, Us=Us, Xdot=Xdot, ang=ang, wrapped=wrapped,
                        minsnap_coefs=np.array([p.coefs for p in ms._polys]),
                        minsnap_get10=ms.get(10.0))


# --------------------------------------------------------------------------
# path A': circular formation (DCF + GVF)
# --------------------------------------------------------------------------
def formation_loop(c_, r, n_ac, t_end, ke, kd, kr, z_des, dt=0.05, X1=None, v_c=15):
    """Loop of 08_CircularFormation_Full.py:73-94 with the per-aircraft centre array of
    09_CircularFormation_diffcentre.py:33,87,103 (script 08 itself crashes for n_ac != 2)."""
    time = np.arange(0, t_end, dt)
    wind = ddg.WindField()
    X1 = np.array([20, 30, -np.pi / 2, 0, 10]) if X1 is None else X1
    X_array = np.zeros((len(time), n_ac, 5)); U_array = np.zeros((len(time), n_ac))
    Ur_array = np.zeros((len(time), n_ac)); e_theta_array = np.zeros((len(time), n_ac - 1))
    R = r * np.ones((n_ac, 1))
    p = np.zeros((2, n_ac))
    B = np.zeros((n_ac, n_ac - 1))
    for i in range(n_ac):
        p[:, i] = X1[:2]; X_array[0, i] = X1
        for j in range(n_ac - 1):
            if i == j: B[i, j] = -1
            elif i == j + 1: B[i, j] = 1
    dcf = ddg.DCFController()
    acs = [ddyn.Aircraft() for _ in range(n_ac)]
    trajs = [ddg.CircleTraj(c_[j, :]) for j in range(n_ac)]
    gvfs = [ddg.GVFcontroller(trajs[j], acs[j], wind) for j in range(n_ac)]
    z_des = np.array(z_des, dtype=float)
    for i in range(1, len(time)):
        t = time[i - 1]
        U_r, e_theta = dcf.get(n_ac, B, c_, p, z_des, kr)
        Rr = U_r + R
        Ur_array[i] = Rr.T; e_theta_array[i] = e_theta.T
        for j in range(n_ac):
            X = X_array[i - 1, j, :]
            e, n, H = trajs[j].get(X, Rr[j])
            U, U1, U2 = gvfs[j].get(X, ke, kd, e, n, H)
            U = np.arctan(U / 9.81)
            U_array[i - 1][j] = U[0] if np.ndim(U) else U
            X_new = acs[j].disc_dyn(X, [U, v_c], wind, t, dt)
            X_array[i][j] = X_new
            p[0][j] = X_new[0]; p[1][j] = X_new[1]
    return X_array, U_array, time, Ur_array, e_theta_array


def golden_formation():
    out = {}
    # C2: 6 aircraft, common centre, RK4 nsub=5
    use_rk4(5)
    n_ac = 6
    X, U, time, Rr, eth = formation_loop(np.zeros((n_ac, 2)), 60, n_ac, 60, 4e-4, 15, 20,
                                         np.ones(n_ac - 1) * (2 * np.pi / n_ac))
    print("C2: X[1,0] =", X[1, 0], "\n    X[1199,0] =", X[1199, 0])
    out.update({"c2/X": X[::4], "c2/U": U[::4], "c2/Rr": Rr[::4], "c2/eth": eth[::4],
                "c2/Xlast": X[-1], "c2/X1": X[1], "c2/T": np.array(len(time))})

    # script 09 run UNMODIFIED (4 aircraft, its own centres, kd=25, z_des=0) under RK4 nsub=5
    s09 = load_script("09_CircularFormation_diffcentre.py", "ref09")
    X9, U9, t9, _, _, Rr9, eth9 = s09.CircularFormationGVF(np.array([0, 0]), 60, 4, 60)
    out.update({"s09/X": X9[::4], "s09/U": U9[::4], "s09/Rr": Rr9[::4], "s09/eth": eth9[::4],
                "s09/Xlast": X9[-1], "s09/T": np.array(len(t9))})
    # my loop restatement must equal script 09 when given script 09's parameters
    c9 = np.array([[0, -20], [25, -40], [25, -80], [0, -100]], dtype=float)
    Xm, Um, _, Rrm, ethm = formation_loop(c9, 60, 4, 60, 4e-4, 25, 20, np.zeros(3))
    assert np.array_equal(Xm, X9) and np.array_equal(Rrm, Rr9), "formation_loop deviates from script 09"

    # states_over_time.csv recipe: script 09, tau_phi=0.9667, LSODA, 200 s  
    use_lsoda()
    _init = ddyn.Aircraft.__init__

    def _init_slow(self):
        _init(self); self.tau_phi = 0.9667
    ddyn.Aircraft.__init__ = _init_slow
    Xc, Uc, tc, _, _, _, _ = s09.CircularFormationGVF(np.array([0, 0]), 60, 4, 200)
    import pandas as pd
    df = pd.read_csv(os.path.join(REF, "src", "states_over_time.csv"))
    csv = np.stack([np.stack([df[f"{s}_{k+1}"].to_numpy() for s in ("x", "y", "psi", "phi", "v")], -1)
                    for k in range(4)], 1)
    err = np.abs(csv - Xc).max()
    print(f"states_over_time.csv reproduced by script 09 (tau_phi=0.9667, LSODA): max abs err {err:.3e}")
    assert err < 1e-9
    # the same recipe under RK4 (nsub=1, then 5): what the CUDA path is compared with, plus its gap to the CSV
    for nsub in (1, 5):
        use_rk4(nsub)
        Xr, Ur, _, _, _, _, _ = s09.CircularFormationGVF(np.array([0, 0]), 60, 4, 200)
        print(f"  RK4 nsub={nsub} vs CSV: max abs {np.abs(Xr - csv).max():.3e}")
        out[f"csv_rk4_{nsub}/X"] = Xr[::20]
        out[f"csv_rk4_{nsub}/Xlast"] = Xr[-1]
    out["csv/X"] = csv[::20]; out["csv/Xlast"] = csv[-1]; out["csv/row1"] = csv[1]
    ddyn.Aircraft.__init__ = _init
    use_rk4(1)

    # single calls
    e, n, H = ddg.CircleTraj(np.array([0, -20])).get(np.array([20, 30, -np.pi / 2, 0, 10]), 60)
    Ug, U1g, U2g = ddg.GVFcontroller(None, None, None).get(np.array([20, 30, -np.pi / 2, 0, 10]), 4e-4, 25, e, n, H)
    out["gvf_known"] = np.array([e, n[0], n[1], Ug, U1g, U2g], dtype=float)
    np.savez_compressed(os.path.join(HERE, "formation.npz"), **out)


# --------------------------------------------------------------------------
# path B: collocation residual / Jacobian (reference EoM, backward Euler) and costs
# --------------------------------------------------------------------------
class SinglePlannerShim:
    """Just the attributes the cost classes read (pattern of test/test_objective.py:183-192)."""
    def __init__(self, N, obj_scale=1.):
        self.num_nodes, self.obj_scale = N, obj_scale
        self._slice_x, self._slice_y, self._slice_psi, self._slice_phi, self._slice_v = \
            [slice(k * N, (k + 1) * N, 1) for k in range(5)]          # 06_optyplan.py:35-39


class MultiPlannerShim:
    def __init__(self, N, n_ac, obj_scale=1.):
        self.num_nodes, self.obj_scale = N, obj_scale
        self.acs = type("acs", (), {"nb_aicraft": n_ac})()
        self._slice_x = [slice((0 + 3 * i) * N, (1 + 3 * i) * N, 1) for i in range(n_ac)]   # 07_multioptyplan.py:41-47
        self._slice_y = [slice((1 + 3 * i) * N, (2 + 3 * i) * N, 1) for i in range(n_ac)]
        self._slice_psi = [slice((2 + 3 * i) * N, (3 + 3 * i) * N, 1) for i in range(n_ac)]
        i_in = N * n_ac * 3
        self._slice_phi = [slice(i_in + i * N, i_in + (1 + i) * N, 1) for i in range(n_ac)]
        i_in += N * n_ac
        self._slice_v = [slice(i_in + i * N, i_in + (1 + i) * N, 1) for i in range(n_ac)]


def discretise(eom, states, inputs, t, h):
    """Backward Euler: xdot -> (x_i - x_p)/h, x -> x_i, u -> u_i  (opty default, confirmed on the
    cached IPOPT solutions).  Returns lambdified residual and dense Jacobian wrt
    [x_i..., x_p..., u_i...] (opty's per-node dense block)."""
    n, q = len(states), len(inputs)
    xi = sym.symbols(f"xi0:{n}"); xp = sym.symbols(f"xp0:{n}"); ui = sym.symbols(f"ui0:{q}")
    sub = {}
    for k, s in enumerate(states):
        sub[s.diff(t)] = (xi[k] - xp[k]) / h
    expr = eom.subs(sub)
    sub2 = {s: xi[k] for k, s in enumerate(states)}
    sub2.update({u: ui[k] for k, u in enumerate(inputs)})
    expr = expr.subs(sub2)
    wrt = list(xi) + list(xp) + list(ui)
    jac = expr.jacobian(wrt)
    f = sym.lambdify(wrt, list(expr), "numpy")
    fj = sym.lambdify(wrt, jac.tolist(), "numpy")
    return f, fj, jac


def opty_sorted_inputs(acs):
    """opty sorts the unknown input trajectories by name."""
    ins = []
    for a in acs:
        ins += [a._sphi(a._st), a._sv(a._st)]
    return sorted(ins, key=lambda s: str(s))


def colloc_eval_reference(acs, eom, t, N, h, free, inst):
    """residual (equation-major), dense per-node Jacobian values (node-major) and instance rows."""
    states = []
    for a in acs: states += list(a._state_symbols)
    inputs = opty_sorted_inputs(acs)
    n, q = len(states), len(inputs)
    f, fj, symjac = discretise(eom, states, inputs, t, h)
    S = free[: n * N].reshape(n, N); Uin = free[n * N:(n + q) * N].reshape(q, N)
    args = [S[k, 1:] for k in range(n)] + [S[k, :-1] for k in range(n)] + [Uin[k, 1:] for k in range(q)]
    res = np.array([np.broadcast_to(r, (N - 1,)) for r in f(*args)])            # (n, N-1)
    J = fj(*args)
    Jd = np.zeros((N - 1, n, 2 * n + q))
    for e in range(n):
        for c in range(2 * n + q):
            Jd[:, e, c] = np.broadcast_to(J[e][c], (N - 1,))
    inst_vals = np.array([free[k * N + node] - val for (k, node, val) in inst])
    residual = np.concatenate([res.reshape(-1), inst_vals])
    # opty's COO structure for the dense blocks
    rows = np.zeros((N - 1, n, 2 * n + q), dtype=np.int64); cols = np.zeros_like(rows)
    for i in range(N - 1):
        r = [e * (N - 1) + i for e in range(n)]
        c = [j * N + i + 1 for j in range(n)] + [j * N + i for j in range(n)] + [n * N + j * N + i + 1 for j in range(q)]
        rows[i] = np.repeat(r, len(c)).reshape(n, -1); cols[i] = np.array(c * n).reshape(n, -1)
    nz = np.array([[symjac[e, c] != 0 for c in range(2 * n + q)] for e in range(n)])
    return residual, Jd, rows, cols, nz


def golden_colloc():
    out = {}
    # planner_timing / triangle known answers
    import io, contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        pt = [d2ou.planner_timing(*a) for a in [(0, 20, 50), (0, 19.98, 50), (0, 9.98, 50), (0, 4.2, 50),
                                                 (0, 5.5, 10), (0, 10, 10), (0, 0.29, 100), (2., 17., 50.)]]
    out["planner_timing"] = np.array(pt, dtype=float)
    tri = d2ou.triangle(np.array([0., 0.]), np.array([50., 0.]), 12., 4.2, 7, go_left=-1)
    out["triangle7"] = np.array(tri)

    # ---- C3: single aircraft, N = 1001, h = 0.02, evaluation point = cached IPOPT solution + noise
    d = np.load(os.path.join(REF, "src", "cache", "optyplan_exp0_1_3.npz"))
    N, h = 1001, 0.02
    sol = np.concatenate([d["sol_x"], d["sol_y"], d["sol_psi"], d["sol_phi"], d["sol_v"]])
    rng = np.random.default_rng(12345)
    noise = np.concatenate([rng.normal(0, s, N) for s in (1., 1., 0.1, 0.05, 0.5)])
    free = sol + noise
    for tag, wind in (("c3", [0., 0.]), ("c3w", [1.5, -2.0])):
        ac = d2ou.Aircraft()
        eom = ac.get_eom(d2ou.WindField(wind))
        inst = [(0, 0, 0.), (1, 0, 0.), (2, 0, 0.), (0, N - 1, 0.), (1, N - 1, 30.), (2, N - 1, np.pi)]
        res, Jd, rows, cols, nz = colloc_eval_reference([ac], eom, ac._st, N, h, free, inst)
        res0, _, _, _, _ = colloc_eval_reference([ac], eom, ac._st, N, h, sol, inst)
        out[f"{tag}/residual"] = res; out[f"{tag}/jac_dense"] = Jd
        out[f"{tag}/residual_at_solution_max"] = np.array(np.abs(res0[:3 * (N - 1)]).max())
        out[f"{tag}/wind"] = np.array(wind)
        print(f"{tag}: |defect| at cached solution = {np.abs(res0[:3*(N-1)]).max():.3e}; nnz/node = {nz.sum()}")
    out["c3/free"] = free; out["c3/sol"] = sol; out["c3/rows"] = rows; out["c3/cols"] = cols; out["c3/nz"] = nz
    out["c3/inst"] = np.array(inst, dtype=float)

    # single-aircraft costs on the cached solution and at the noisy point
    P = SinglePlannerShim(N, 1.)
    costs = {
        "airvel": d2ou.CostAirVel(12.),
        "bank": d2ou.CostBank(),
        "input": d2ou.CostInput(12., 1., 50.),
        "obs0": d2ou.CostObstacle((30, 0), 15., kind=0),
        "obs1": d2ou.CostObstacle((5, 15), 10., kind=1),
        "composit": d2ou.CostComposit(((5, 15, 10),), vsp=15., kobs=.5, kvel=.5, kbank=1.),
        "composit1": d2ou.CostComposit(((5, 15, 10), (-3., 20., 6.)), vsp=12., kobs=2., kvel=.7, kbank=1.5, obs_kind=1),
    }
    for name, c in costs.items():
        for tag, fr in (("sol", sol), ("noisy", free)):
            out[f"cost1/{name}/{tag}/cost"] = np.array(c.cost(fr, P))
            out[f"cost1/{name}/{tag}/grad"] = np.asarray(c.cost_grad(fr, P))
        print(f"cost {name}: {float(out[f'cost1/{name}/sol/cost']):.15e} |grad| {np.linalg.norm(out[f'cost1/{name}/sol/grad']):.15e}")
    bank_max = d2ou.CostBank(); bank_max.use_mean = False        # max-squared-bank mode, opty_utils.py:72,78-81
    for tag, fr in (("sol", sol), ("noisy", free)):
        out[f"cost1/bankmax/{tag}/cost"] = np.array(bank_max.cost(fr, SinglePlannerShim(N, 2.5)))
        out[f"cost1/bankmax/{tag}/grad"] = np.asarray(bank_max.cost_grad(fr, SinglePlannerShim(N, 2.5)))
    P2 = SinglePlannerShim(N, 3.5)
    out["cost1/input_scaled/noisy/cost"] = np.array(costs["input"].cost(free, P2))
    out["cost1/input_scaled/noisy/grad"] = costs["input"].cost_grad(free, P2)

    # ---- small multi-aircraft case with opty-dense structure: n_ac = 3, N = 20, h = 0.1, wind
    for tag, n_ac, N, h, wind in (("m3", 3, 20, 0.1, [0.5, -1.0]), ("c4", 16, 500, 0.02, [0., 0.])):
        acs = d2mou.AircraftSet(n_ac)
        eom = acs.get_eom(d2ou.WindField(wind))
        P = MultiPlannerShim(N, n_ac, 1.)
        ang = 2 * np.pi * np.arange(n_ac) / n_ac
        p0s = np.stack([100 * np.cos(ang), 100 * np.sin(ang), ang + np.pi], 1)
        p1s = np.stack([-100 * np.cos(ang), -100 * np.sin(ang), ang + np.pi], 1)
        duration = (N - 1) * h
        freeM = np.zeros(5 * n_ac * N)
        for i in range(n_ac):
            ig = d2ou.triangle(p0s[i, :2], p1s[i, :2], 12., duration, N, go_left=-1.)     # 07_multioptyplan.py:108
            freeM[P._slice_x[i]], freeM[P._slice_y[i]], freeM[P._slice_psi[i]], freeM[P._slice_phi[i]], freeM[P._slice_v[i]] = ig
        rng = np.random.default_rng(12345)
        sig = np.zeros_like(freeM)
        for i in range(n_ac):
            sig[P._slice_x[i]] = 1.; sig[P._slice_y[i]] = 1.; sig[P._slice_psi[i]] = 0.1
            sig[P._slice_phi[i]] = 0.05; sig[P._slice_v[i]] = 0.5
        freeM = freeM + rng.normal(0., 1., freeM.size) * sig
        # instance constraints in the planner's order (07_multioptyplan.py:53-56): all t0 then all t1
        inst = []
        for i in range(n_ac):
            inst += [(3 * i + k, 0, p0s[i, k]) for k in range(3)]
        for i in range(n_ac):
            inst += [(3 * i + k, N - 1, p1s[i, k]) for k in range(3)]
        # NOTE the planner's free-vector convention for inputs is numeric order (phi_0..phi_{n-1}, v_0..),
        # opty's is name-sorted; for n_ac <= 10 they coincide.  The golden uses NUMERIC order (the planner's
        # slices, which every cost class reads); the engine takes an explicit permutation.
        states = []
        for a in acs.aircraft: states += list(a._state_symbols)
        inputs = [a._sphi(a._st) for a in acs.aircraft] + [a._sv(a._st) for a in acs.aircraft]
        n, q = len(states), len(inputs)
        f, fj, symjac = discretise(eom, states, inputs, acs.st, h)
        S = freeM[: n * N].reshape(n, N); Uin = freeM[n * N:].reshape(q, N)
        args = [S[k, 1:] for k in range(n)] + [S[k, :-1] for k in range(n)] + [Uin[k, 1:] for k in range(q)]
        res = np.array([np.broadcast_to(r, (N - 1,)) for r in f(*args)]).reshape(-1)
        inst_vals = np.array([freeM[k * N + node] - val for (k, node, val) in inst])
        out[f"{tag}/free"] = freeM; out[f"{tag}/residual"] = np.concatenate([res, inst_vals])
        out[f"{tag}/inst"] = np.array(inst, dtype=float); out[f"{tag}/wind"] = np.array(wind)
        out[f"{tag}/p0s"] = p0s; out[f"{tag}/p1s"] = p1s
        J = fj(*args)
        nzmask = np.array([[symjac[e, c] != 0 for c in range(2 * n + q)] for e in range(n)])
        ee, cc = np.nonzero(nzmask)
        out[f"{tag}/nz_eq"] = ee; out[f"{tag}/nz_col"] = cc
        out[f"{tag}/jac_nz"] = np.stack([np.broadcast_to(J[e][c], (N - 1,)) for e, c in zip(ee, cc)], 1)   # (N-1, nnz/node)
        if tag == "m3":
            Jd = np.zeros((N - 1, n, 2 * n + q))
            for e in range(n):
                for c in range(2 * n + q):
                    Jd[:, e, c] = np.broadcast_to(J[e][c], (N - 1,))
            out[f"{tag}/jac_dense"] = Jd
            # name-sorted variant for 12 aircraft: the permutation opty would apply
        print(f"{tag}: n_ac={n_ac} N={N} nnz/node={nzmask.sum()} residual max {np.abs(res).max():.3e}")
        # multi-aircraft costs (reference classes): CostInput, obstacles on aircraft 0, collision (0,1)
        mc = {
            "input": d2mou.CostInput(vsp=12., kv=70., kphi=1.),
            "airvel": d2mou.CostAirvel(12.),
            "bank": d2mou.CostBank(),
            "obs0": d2mou.CostObstacle((60., 5.), 12., kind=0),
            "obs1": d2mou.CostObstacle((60., 5.), 12., kind=1),
            "collision": d2mou.CostCollision(r=10., k=2.),
            "composit": d2mou.CostComposit(kvel=70., kbank=1., kobs=0.5, kcol=10., vsp=12.,
                                           obss=((60., 5., 12.), (-20., 30., 8.)), obs_kind=1, rcol=10.),
            "composit_nocol": d2mou.CostComposit(kvel=2., kbank=1., vsp=12.),
        }
        for name, c in mc.items():
            out[f"{tag}/cost/{name}/cost"] = np.array(c.cost(freeM, P))
            out[f"{tag}/cost/{name}/grad"] = np.asarray(c.cost_grad(freeM, P)) if tag == "m3" or name in ("composit", "collision") \
                else np.asarray(c.cost_grad(freeM, P))[::7]
    # opty's name-sorted input order for 12 aircraft
    acs12 = d2mou.AircraftSet(12)
    names = [str(s) for s in opty_sorted_inputs(acs12.aircraft)]
    out["sorted_inputs_12"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "colloc.npz"), **out)

    # backward-Euler defect of every shipped solution (pins the residual convention)
    defects = {}
    cache = os.path.join(REF, "src", "cache")
    for fn in sorted(os.listdir(cache)) + ["../sample.npz"]:
        dd = np.load(os.path.join(cache, fn))
        tt = dd["sol_time"]; hh = tt[1] - tt[0]; w = dd["wind"][0]
        x, y, psi, phi, v = (dd[k] for k in ("sol_x", "sol_y", "sol_psi", "sol_phi", "sol_v"))
        r1 = (x[1:] - x[:-1]) / hh - v[1:] * np.cos(psi[1:]) + w[0]
        r2 = (y[1:] - y[:-1]) / hh - v[1:] * np.sin(psi[1:]) + w[1]
        r3 = (psi[1:] - psi[:-1]) / hh - 9.81 / v[1:] * np.tan(phi[1:])
        defects[fn] = max(np.abs(r1).max(), np.abs(r2).max(), np.abs(r3).max())
        print(f"  defect {fn}: N={len(tt)} h={hh:.3f} max={defects[fn]:.3e}")


def golden_tabulated(sim):
    """TrajTabulated (d2d/trajectory_factory.py:149-171) on two planner solutions shipped with the reference,
    tracked by the DFFF controller: the hand-over from path B's output to path A's input."""
    import io, contextlib
    use_rk4(1)
    out = {}
    for tag, fn in (("exp0", "optyplan_exp0.npz"), ("exp13", "optyplan_exp13 - some traj.npz")):
        path = os.path.join(REF, "src", "cache", fn)
        with contextlib.redirect_stdout(io.StringIO()):
            traj = ddtf.TrajTabulated(path)
        d = np.load(path)
        for k in ("sol_time", "sol_x", "sol_y", "sol_psi", "sol_phi", "sol_v", "wind"):
            out[f"{tag}/{k}"] = d[k]
        time = np.arange(0., traj.duration + 0.5, 0.01)          # runs past the end: exercises the wrap to row 0
        wind = ddg.WindField([0.5, -0.3])
        ac = ddyn.Aircraft()
        X0 = ddg.DiffFlatness.state_and_input_from_output(traj.get(0.), wind.sample(0, None), ac)[0] + np.array([1., -2., 0.1, 0., 0.3])
        X, U, Yref, Xref, K = run_dfff(sim, time, traj, wind, X0, np.zeros((len(time), 5)))
        out[f"{tag}/X"] = X[::5]; out[f"{tag}/U"] = U[::5]; out[f"{tag}/Yref"] = Yref[::7]; out[f"{tag}/X0"] = X0
        out[f"{tag}/T"] = np.array(len(time)); out[f"{tag}/Xlast"] = X[-1]
        print(f"tabulated {tag}: {len(d['sol_time'])} rows, T={len(time)}, X[-1]={X[-1]}")
    np.savez_compressed(os.path.join(HERE, "tabulated.npz"), **out)


def golden_tracker():
    """Controllers.DiffController (5-state LQR, Controllers.py:139-186) driven by implement_controller of
    10_opt_traj_tracking.py:27-90 on two planner outputs shipped with the reference.  The script's own
    loop is used unmodified; only Aircraft.disc_dyn is the fixed-step RK4 (nsub = 10: dt = 0.1 s, tau_phi = 0.01 s)."""
    import pandas as pd
    cwd = os.getcwd()
    os.chdir(os.path.join(REF, "src"))
    use_rk4(10)
    try:
        s10 = load_script("10_opt_traj_tracking.py", "ref10")       # runs its main() once on import (plots are stubbed)
        import Controllers as ctl_mod
        out = {}
        X0s = ((0, 40, np.deg2rad(0), 0, 12), (25, 20, np.deg2rad(0), 0, 12), (25, -20, np.deg2rad(0), 0, 12), (0, -40, np.deg2rad(0), 0, 12))
        for tag, fn, w in (("simple", "opt_states_simple_traj.csv", [0, 0]), ("opt", "opt_states.csv", [1.0, -0.5]),
                           ("hf", "opt_states_hf.csv", [0, 0]), ("stline", "opt_states_st_line.csv", [0.5, 0.5]), ("inf", "inf_traj_10s.csv", [0, 0])):
            df = pd.read_csv(fn)
            # capture the gains: implement_controller creates its own DiffController, so record through the class
            Ks = []
            orig = ctl_mod.DiffController.ComputeGain

            def rec(self, *a, _o=orig, **k):
                r = _o(self, *a, **k); Ks.append(self.K[-1].copy()); return r
            ctl_mod.DiffController.ComputeGain = rec
            X0s_run = X0s if tag in ("simple", "opt") else tuple((df[f"x_{i+1}"].iloc[0] + 1., df[f"y_{i+1}"].iloc[0] - 1., df[f"psi_{i+1}"].iloc[0], 0., 12.) for i in range(4))
            X, U, Xr, Yd, Ydd, t, dX = s10.implement_controller(4, df, 10, w, X0s_run)
            ctl_mod.DiffController.ComputeGain = orig
            out[f"{tag}/time"] = t; out[f"{tag}/wind"] = np.array(w, dtype=float); out[f"{tag}/X0s"] = np.array(X0s_run, dtype=float)
            out[f"{tag}/x_ref"] = np.stack([df[f"x_{i+1}"].to_numpy() for i in range(4)], 1)
            out[f"{tag}/y_ref"] = np.stack([df[f"y_{i+1}"].to_numpy() for i in range(4)], 1)
            out[f"{tag}/X"] = X; out[f"{tag}/U"] = U
            if tag in ("simple", "opt"):
                out[f"{tag}/Xr"] = Xr; out[f"{tag}/dX"] = dX; out[f"{tag}/Yd"] = Yd; out[f"{tag}/Ydd"] = Ydd
                out[f"{tag}/K"] = np.array(Ks).reshape(len(t) - 1, 4, 2, 5)
            print(f"tracker {tag}: T={len(t)} X[-1,0]={X[-1,0]} end-point miss of aircraft 1: "
                  f"{np.hypot(X[-1,0,0]-df['x_1'].iloc[-1], X[-1,0,1]-df['y_1'].iloc[-1]):.3f} m")
        # single calls of ComputeFlatness with a non-zero third derivative
        rng = np.random.default_rng(5)
        n = 32
        Y = rng.normal(0, 1, (n, 4, 2)) * np.array([50., 8., 2., 0.5])[None, :, None]
        W = rng.normal(0, 2, (n, 2))
        Xr = np.zeros((n, 5)); Ur = np.zeros((n, 2))
        for i in range(n):
            Xr[i], Ur[i] = ctl_mod.DiffFlatness(list(W[i])).ComputeFlatness(0., Y[i, 0], Y[i, 1], Y[i, 2], Y[i, 3])
        out["flat/Y"] = Y; out["flat/W"] = W; out["flat/Xr"] = Xr; out["flat/Ur"] = Ur
        np.savez_compressed(os.path.join(HERE, "tracker.npz"), **out)
    finally:
        os.chdir(cwd)
        use_rk4(1)


def golden_pursuit(sim):
    """PurePursuitControler (d2d/guidance.py:204-245) in the run_simulation loop, on the square patrol (composite of lines
    and arcs, 3.3 k path samples) and on a circle; wind on the second case.  RK4 nsub 1 stands in for LSODA as everywhere."""
    out = {}
    use_rk4(1)
    for tag, traj, w, X0, T in (("square", ddtf.TrajSquare(), [0., 0.], [5., -3., 0.3, 0., 10.], 1500),
                                ("circle", ddtf.TrajCircle(), [2., -1.], [25., 5., 1.2, 0., 9.], 1200)):
        ctl = ddg.PurePursuitControler(traj)
        ac, wind = ddyn.Aircraft(), ddg.WindField(w)
        time = np.arange(T) * 0.01
        X, U, _ = sim.run_simulation(time, ac, wind, ctl, np.array(X0, dtype=float), np.zeros((T, 5)))
        idx = np.array([int(np.argmin(np.linalg.norm(ctl.pts_2d - r, axis=1))) for r in np.array(ctl.ref_pos)])
        out[f"{tag}/pts"], out[f"{tag}/time"], out[f"{tag}/wind"], out[f"{tag}/X0"] = ctl.pts_2d, time, np.array(w), np.array(X0)
        out[f"{tag}/X"], out[f"{tag}/U"], out[f"{tag}/idx"], out[f"{tag}/carrot"] = X, U, idx, np.array(ctl.carrot)
        print(f"pursuit {tag}: {len(ctl.pts_2d)} path samples, final state {X[-1]}")
    np.savez_compressed(os.path.join(HERE, "pursuit.npz"), **out)


def golden_spline(sim):
    """TrajSpline (d2d/trajectory_factory.py:189-211, default way points: the only ones its constructor accepts) sampled
    over two periods, SplineOne samples, and the DFFF closed loop on it."""
    traj = ddtf.TrajSpline()
    ts = np.concatenate([np.linspace(0., 2.2 * traj.duration, 400), [0., traj.duration * 0.5, traj.duration - 1e-9]])
    Y = np.array([traj.get(t) for t in ts])
    xs = np.arange(0, 30.5, 0.5); ys = np.sin(xs / 30. * np.pi / 2)
    one = ddtf.SplineOne(xs, ys)
    t1 = np.linspace(0., 30., 77)
    Y1 = np.array([one.get(t) for t in t1])
    use_rk4(1)
    time = np.arange(0, 20, 0.01)
    wind = ddg.WindField([1., 0.5])
    ac = ddyn.Aircraft()
    X0 = ddg.DiffFlatness.state_and_input_from_output(traj.get(0.), wind.sample(0, None), ac)[0] + np.array([2., -2., 0.1, 0., 0.])
    X, U, Yref, Xref, K = run_dfff(sim, time, traj, wind, X0, np.zeros((len(time), 5)))
    print("spline: duration", traj.duration, "X[-1]", X[-1])
    # SpaceIndexedTraj beyond TrajSiDemo: circle geometry (TrajSiSpline's, trajectory_factory.py:246-247) and the scalar
    # dynamics of d2d/trajectory.py:13-38
    si = {}
    circ = ddt.TrajectoryCircle(c=[30., 30.], r=30., v=2 * np.pi * 30., t0=0., alpha0=0, dalpha=3 * np.pi / 2)
    line = ddt.TrajectoryLine([0., 0.], [80., 40.], v=np.hypot(80., 40.), t0=0.)
    cases = {"circle_affine": (circ, ddt.AffineOne(1. / 30., 0., duration=30.)),
             "circle_poly": (circ, ddt.PolynomialOne([0, 0, 0, 0], [1, 0, 0, 0], 25.)),
             "line_sin": (line, ddt.SinOne(c=0.5, a=0.6, om=0.4, duration=2 * np.pi / 0.4)),
             "circle_sin": (circ, ddt.SinOne(c=0.4, a=0.3, om=0.7, duration=2 * np.pi / 0.7))}
    for tag, (geom, dyn) in cases.items():
        tr = ddt.SpaceIndexedTraj(geom, dyn)
        tsi = np.linspace(0., tr.duration, 200)
        si[f"si/{tag}/t"], si[f"si/{tag}/Y"] = tsi, np.array([tr.get(t) for t in tsi])
    tr = ddt.SpaceIndexedTraj(*cases["circle_poly"])
    time2 = np.arange(0, 25, 0.01)
    wind2 = ddg.WindField([0.5, -0.5])
    X0b = ddg.DiffFlatness.state_and_input_from_output(tr.get(0.5), wind2.sample(0, None), ac)[0] + np.array([1., 1., 0., 0., 1.])
    Xb, Ub, _, _, _ = run_dfff(sim, time2[50:], tr, wind2, X0b, np.zeros((len(time2) - 50, 5)))
    si["si/run/time"], si["si/run/wind"], si["si/run/X0"], si["si/run/X"], si["si/run/U"] = time2[50:], np.array([0.5, -0.5]), X0b, Xb, Ub
    # TrajSiSpline (trajectory_factory.py:241-285) as ScenCircle builds it by default (scenario.py:109): the constructor's
    # optimiser leaves the object with the last FooOne it probed; its knots are stored so that the engine-side class can be
    # given the same dynamic (a BFGS path is not bit-reproducible across builds), then the whole default "circle" scenario
    import d2d.scenario as dds
    scen = dds.ScenCircle()
    tr_s = scen.trajs[0]
    t_s = np.linspace(0., tr_s.duration - 0.01, 300)
    si["sisp/xs"], si["sisp/ys"], si["sisp/duration"] = np.asarray(tr_s._dyn.xs, float), np.asarray(tr_s._dyn.ys, float), np.array(tr_s.duration)
    si["sisp/t"], si["sisp/Y"] = t_s, np.array([tr_s.get(t) for t in t_s])
    Xs, Us, _, Xrs, Ks = run_dfff(sim, scen.time, tr_s, scen.windfield, scen.X0s[0], scen.perts[0])
    si["sisp/time"], si["sisp/X0"], si["sisp/X"], si["sisp/U"] = scen.time, np.asarray(scen.X0s[0], float), Xs[::10], Us[::10]
    vair = np.linalg.norm(si["sisp/Y"][:, 1] - np.array([5., 0.]), axis=1)
    print("sispline: duration", tr_s.duration, "knots ys", si["sisp/ys"], "air speed range", vair.min(), vair.max(), "X[-1]", Xs[-1])
    np.savez_compressed(os.path.join(HERE, "spline.npz"), duration=traj.duration, ts=ts, Y=Y, one_xs=xs, one_ys=ys, one_t=t1, one_Y=Y1,
                        time=time, wind=np.array([1., 0.5]), X0=X0, X=X, U=U, **si)


def main():
    what = sys.argv[1:] or ["c1", "scen", "units", "form", "colloc", "tab", "tracker", "pursuit", "spline"]
    sim = load_script("05_test_simulation.py", "ref05")
    if "c1" in what: golden_c1(sim)
    if "scen" in what: golden_scenarios(sim)
    if "units" in what: golden_units()
    if "form" in what: golden_formation()
    if "colloc" in what: golden_colloc()
    if "tab" in what: golden_tabulated(sim)
    if "tracker" in what: golden_tracker()
    if "pursuit" in what: golden_pursuit(sim)
    if "spline" in what: golden_spline(sim)


if __name__ == "__main__":
    main()
