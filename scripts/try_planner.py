"""Solves exp_0-like planner problems on the GPU and checks them with the collocation constraints."""
import sys, time, numpy as np
sys.path.insert(0, "drone-sim-python_b200")
import torch
from d2d_b200 import planner as pl, opty_utils as d2ou, multiopty_utils as d2mou

def report(p, info, t):
    res = p.prob.con(p.solution)
    print(f"  its {info['iterations']} nfev {info['nfev']} outer {info['outer']}  time {t:.2f}s  cost {p.prob.obj(p.solution):.6e}  "
          f"|con| {np.abs(res).max():.2e}  feasible {info['feasible']}")

class exp_c3(pl.exp_0):
    t1, hz = 20., 50.
for exp in (pl.exp_0, exp_c3):
    p = pl.Planner(exp)
    p.configure(tol=1e-8)
    t0 = time.time(); info = p.run(verbose="-v" in sys.argv); torch.cuda.synchronize(); report(p, info, time.time() - t0)
    t0 = time.time(); info = p.run(n_starts=64); torch.cuda.synchronize(); report(p, info, time.time() - t0)
    print("   multi-start costs", np.sort(info["cost"])[:5], "feasible", (info["c_max"] < 1e-6).sum())

class scen4:
    t0, t1, hz = 0., 10., 20.
    p0s = [(0., 0., 0.), (0., 30., 0.), (80., 0., np.pi), (80., 30., np.pi)]
    p1s = [(80., 30., 0.), (80., 0., 0.), (0., 30., np.pi), (0., 0., np.pi)]
    vref, obj_scale = 12., 1.
    wind = d2ou.WindField(w=[0., 0.])
    cost = d2mou.CostComposit(kvel=1., kbank=1., kcol=10., vsp=12., rcol=5., all_pairs=True, obss=[(40., 15., 6.)], kobs=5., obs_kind=1)
    phi_constraint = (-np.deg2rad(40.), np.deg2rad(40.))
    v_constraint = (9., 15.)
    x_constraint = y_constraint = None
p = pl.MultiPlanner(scen4)
p.configure(tol=1e-8)
for ns in (1, 32):
    t0 = time.time(); info = p.run(n_starts=ns, verbose="-v" in sys.argv); torch.cuda.synchronize(); report(p, info, time.time() - t0)
    print("   costs", np.sort(info["cost"])[:5], "feasible", (info["c_max"] < 1e-6).sum())
p.interpret_solution()
d = min(np.hypot(p.sol_x[a] - p.sol_x[b], p.sol_y[a] - p.sol_y[b]).min() for a in range(4) for b in range(a))
print("   min separation", d, " min distance of aircraft 0 to obstacle", np.hypot(p.sol_x[0] - 40., p.sol_y[0] - 15.).min())
