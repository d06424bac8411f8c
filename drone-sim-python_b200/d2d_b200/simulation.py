"""Closed-loop simulation on the engine: `run_simulation` / `test_simulation` of 05_test_simulation.py and
`CircularFormationGVF` of 08_CircularFormation_Full.py / 09_CircularFormation_diffcentre.py, plus the batched
forms (`rollout`, `formation_rollout`) that the single-scenario functions are thin views of.

Host arrays in, host arrays out: each call copies its inputs to HBM, launches the fused rollout kernel and
copies the requested logs back."""
import numpy as np
import torch

from . import trajectory as ddt
from .dynamic import Aircraft
from .engine import get_engine
from .guidance import DFFFController, WindField


def _perts_csr(eng, perts_list, T):
    """Dense per-scenario perturbation arrays (T,5) -> sparse CSR events on the device (or None)."""
    begin, steps, dxs = [0], [], []
    for p in perts_list:
        if p is not None:
            p = np.asarray(p, dtype=np.float64)
            rows = np.nonzero(np.any(p != 0., axis=1))[0]
            rows = rows[rows >= 1]                           # perts[0] is never applied (05_test_simulation.py:28-32)
            steps += list(rows); dxs += [p[r] for r in rows]
        begin.append(len(steps))
    if not steps:
        return None
    return (eng.to_device(np.asarray(begin, np.int32)), eng.to_device(np.asarray(steps, np.int32)),
            eng.to_device(np.ascontiguousarray(np.asarray(dxs, np.float64).T)))


class RolloutResult:
    """X (B,T_log,5), U (B,T_log,2), optional Xref (B,T_log,5), K (B,T_log,2,5), X_final (B,5),
    sum_sq_err (B,), max_err (B,), flags (B,)."""
    pass


def rollout(time, trajs, wind, X0, perts=None, tau_phi=0.01, tau_v=1., nsub=1, log_every=1, log_ref=False,
            final_control=True, gains=None, engine=None, chunk_steps=None, return_log=True):
    """Batched run_simulation (05_test_simulation.py:21-34) for B aircraft-scenarios.

    time   (T,) sample grid shared by all scenarios (np.arange as the reference builds it)
    trajs  list of B Trajectory objects, or a CircleBatch / MinSnapBatch
    wind   (2,) or (B,2);  X0 (B,5);  perts None or list of B arrays (T,5) / None
    tau_phi, tau_v scalars or (B,) arrays;  nsub RK4 sub-steps per control step
    """
    eng = engine or get_engine()
    time = np.ascontiguousarray(time, dtype=np.float64)
    T = len(time)
    X0 = np.asarray(X0, dtype=np.float64).reshape(-1, 5)
    B = len(X0)
    packed = ddt.pack(trajs)
    if packed.n_traj != B:
        raise ValueError(f"{packed.n_traj} trajectories for {B} initial states")
    tab = eng.table(packed)
    wind = np.broadcast_to(np.asarray(wind, dtype=np.float64).reshape(-1, 2), (B, 2))
    ac = np.stack([np.broadcast_to(np.asarray(tau_phi, dtype=np.float64), (B,)), np.broadcast_to(np.asarray(tau_v, dtype=np.float64), (B,))])
    X0d, Wd, acd, td = eng.to_device(np.ascontiguousarray(X0.T)), eng.to_device(np.ascontiguousarray(wind.T)), eng.to_device(ac), eng.to_device(time)
    pd = _perts_csr(eng, perts, T) if perts is not None else None
    n_rows = (T - 1) // log_every + 1
    X_log = eng.empty(n_rows, 5, B) if return_log else None
    U_log = eng.empty(n_rows, 2, B) if return_log else None
    Xr_log = eng.empty(n_rows, 5, B) if (log_ref and return_log) else None
    K_log = eng.empty(n_rows, 6, B) if (log_ref and return_log) else None
    sum_sq, max_err = eng.zeros(B), eng.zeros(B)
    flags = eng.zeros(B, dtype=torch.int32)
    care = eng.zeros(3, B)
    pop = eng.zeros(2)
    X_final = eng.empty(5, B)
    chunk = chunk_steps or (T - 1)
    i, Xc = 0, X0d
    while True:
        j = min(i + chunk, T - 1)
        last = j == T - 1
        eng.rollout_dfff(tab, Xc, Wd, acd, td, i, j, nsub=nsub, final_control=(final_control and last), gains=gains,
                         log_every=log_every, X_log=X_log, U_log=U_log, Xr_log=Xr_log, K_log=K_log, X_final=X_final,
                         sum_sq_err=sum_sq, max_err=max_err, flags=flags, care_state=care, pop_stats=pop, perts=pd)
        if last:
            break
        i, Xc = j, X_final
    res = RolloutResult()
    res.time_log = time[::log_every]
    res.X_final = X_final.cpu().numpy().T
    res.sum_sq_err, res.max_err, res.flags = sum_sq.cpu().numpy(), max_err.cpu().numpy(), flags.cpu().numpy()
    res.pop_sum_sq_err, res.pop_max_err = (float(v) for v in pop.cpu().numpy())
    if return_log:
        res.X = X_log.permute(2, 0, 1).contiguous().cpu().numpy()
        res.U = U_log.permute(2, 0, 1).contiguous().cpu().numpy()
        if log_ref:
            res.Xref = Xr_log.permute(2, 0, 1).contiguous().cpu().numpy()
            K3 = K_log.permute(2, 0, 1).contiguous().cpu().numpy().reshape(B, n_rows, 2, 3)
            res.K = np.concatenate([K3, np.zeros((B, n_rows, 2, 2))], axis=3)
    return res


def run_simulation(time, aircraft, windfield, ctl, X0, perts):
    """Drop-in for run_simulation(time, aircraft, windfield, ctl, X0, perts) -> X (T,5), U (T,2), Yref (T,4,2)
    (05_test_simulation.py:21-34).  `ctl` must be a DFFFController; its Xref / K logs are filled as upstream."""
    if not isinstance(ctl, DFFFController):
        raise TypeError("the engine implements the DFFF controller (scen.ppctl is False in every reference scenario)")
    time = np.asarray(time, dtype=np.float64)
    W = windfield.sample(time[0], None)
    res = rollout(time, [ctl.traj], W, np.asarray(X0, dtype=np.float64).reshape(1, 5), perts=[perts],
                  tau_phi=aircraft.tau_phi, tau_v=aircraft.tau_v, nsub=getattr(aircraft, "nsub", 1), log_ref=True)
    ctl.Xref += list(res.Xref[0])
    ctl.K += list(res.K[0])
    Yref = ctl.traj.get_many(time)
    return res.X[0], res.U[0], Yref


def test_simulation(scen, **_ignored):
    """Every aircraft of a scenario in ONE launch (the loop of 05_test_simulation.py:37-53).
    Returns Xs, Us, Yrefs lists like the upstream loop accumulates."""
    n = len(scen.trajs)
    W = scen.windfield.sample(scen.time[0], None)
    res = rollout(scen.time, scen.trajs, W, np.asarray([np.asarray(x, dtype=np.float64) for x in scen.X0s[:n]]),
                  perts=list(scen.perts), tau_phi=[a.tau_phi for a in scen.aircrafts], tau_v=[a.tau_v for a in scen.aircrafts])
    Yrefs = [tr.get_many(scen.time) for tr in scen.trajs]
    return list(res.X), list(res.U), Yrefs


def chain_incidence(n_ac):
    """Incidence matrix of the line graph 0-1-...-(n_ac-1) (08_CircularFormation_Full.py:49-60)."""
    B = np.zeros((n_ac, max(n_ac - 1, 0)))
    for j in range(n_ac - 1):
        B[j, j], B[j + 1, j] = -1, 1
    return B


def formation_rollout(c, r, n_ac, time_len, dt, ke, kd, kr, z_des, X0, v_c=15., nsub=5, tau_phi=0.01, tau_v=1.,
                      B=None, log_every=1, engine=None, return_log=True):
    """F formations of n_ac aircraft in one launch.  c (F,n_ac,2), r scalar or (F,n_ac), X0 (F,n_ac,5).
    Returns dict with X (F,T,n_ac,5), U (F,T,n_ac), Rr (F,T,n_ac), e_theta (F,T,n_e) in the reference's row
    conventions (08_CircularFormation_Full.py:77-78,86: row i of Rr / e_theta belongs to step i-1 -> i)."""
    eng = engine or get_engine()
    c = np.asarray(c, dtype=np.float64).reshape(-1, n_ac, 2)
    F = len(c)
    M = F * n_ac
    X0 = np.broadcast_to(np.asarray(X0, dtype=np.float64), (F, n_ac, 5)).reshape(M, 5)
    r = np.broadcast_to(np.asarray(r, dtype=np.float64), (F, n_ac)).reshape(M)
    Binc = chain_incidence(n_ac) if B is None else np.asarray(B, dtype=np.float64)
    n_e = Binc.shape[1]
    ac = np.stack([np.full(M, tau_phi, dtype=np.float64), np.full(M, tau_v, dtype=np.float64)])
    T = int(time_len)
    n_rows = (T - 1) // log_every + 1
    X_log = eng.empty(n_rows, 5, M) if return_log else None
    U_log = eng.zeros(n_rows, M) if return_log else None
    Rr_log = eng.zeros(n_rows, M) if return_log else None
    eth_log = eng.zeros(n_rows, F * max(n_e, 1)) if return_log else None
    flags = eng.zeros(M, dtype=torch.int32)
    Xf = eng.rollout_formation(n_ac, Binc, z_des, eng.to_device(np.ascontiguousarray(X0.T)),
                               eng.to_device(np.ascontiguousarray(c.reshape(M, 2).T)), eng.to_device(r), eng.to_device(ac),
                               ke, kd, kr, v_c, dt, 0, T - 1, nsub, log_every=log_every, X_log=X_log, U_log=U_log,
                               Rr_log=Rr_log, eth_log=eth_log, flags=flags)
    out = {"X_final": Xf.cpu().numpy().T.reshape(F, n_ac, 5), "flags": flags.cpu().numpy().reshape(F, n_ac)}
    if return_log:
        out["X"] = X_log.cpu().numpy().reshape(n_rows, 5, F, n_ac).transpose(2, 0, 3, 1)
        out["U"] = U_log.cpu().numpy().reshape(n_rows, F, n_ac).transpose(1, 0, 2)
        if log_every == 1:                                   # reference row convention: row i holds step (i-1)'s value
            Rr = Rr_log.cpu().numpy().reshape(n_rows, F, n_ac).transpose(1, 0, 2)
            eth = eth_log.cpu().numpy()[:, :F * n_e].reshape(n_rows, F, n_e).transpose(1, 0, 2)
            out["Rr"] = np.concatenate([np.zeros((F, 1, n_ac)), Rr[:, :-1]], axis=1)
            out["e_theta"] = np.concatenate([np.zeros((F, 1, n_e)), eth[:, :-1]], axis=1)
    return out


def CircularFormationGVF(c, r, n_ac, t_end, ke=0.0004, kd=15, kr=20, z_des=None, t_step=0.05, nsub=5, tau_phi=0.01):
    """Drop-in for CircularFormationGVF(c, r, n_ac, t_end) (08_CircularFormation_Full.py:21-97) ->
    X_array, U_array, time, U1_array, U2_array, Ur_array, e_theta_array.  `c` may be one centre (2,) or one per
    aircraft (n_ac,2) as in script 09.  U1/U2 (debug split of the GVF output) are not logged by the rollout
    kernel and are returned as zeros; use GVFcontroller.get for them."""
    time = np.arange(0, t_end, t_step)
    c = np.asarray(c, dtype=np.float64)
    c_ = np.ones((n_ac, 2)) * c if c.ndim == 1 else c
    z_des = np.ones(n_ac - 1) * (np.pi * 2 / n_ac) if z_des is None else np.asarray(z_des, dtype=np.float64)
    X1 = np.array([20, 30, -np.pi / 2, 0, 10])               # 08_CircularFormation_Full.py:44
    o = formation_rollout(c_[None], r, n_ac, len(time), t_step, ke, kd, kr, z_des, X1, nsub=nsub, tau_phi=tau_phi)
    zeros = np.zeros((len(time), n_ac))
    return o["X"][0], o["U"][0], time, zeros, zeros.copy(), o["Rr"][0], o["e_theta"][0]
