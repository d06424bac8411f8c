import sys, time, numpy as np
sys.path.insert(0, "drone-sim-python_b200")
import torch
from d2d_b200 import planner as pl, multiopty_scenarios as S
names = sys.argv[1:] or ["exp_0", "exp_1", "exp_1_0", "exp_2", "exp_3", "exp_4", "exp_5", "exp_5_1", "gvf_trial_3ac", "inf_traj_4ac"]
for nm in names:
    exp = getattr(S, nm)
    for case in range(min(exp.ncases, 2)):
        exp.set_case(case)
        p = pl.MultiPlanner(exp)
        p.configure(tol=exp.tol, max_iter=exp.max_iter)
        t0 = time.time(); info = p.run(initial_guess=p.get_initial_guess("tri"), n_starts=8); torch.cuda.synchronize(); dt = time.time() - t0
        res = np.abs(p.prob.con(p.solution)).max()
        n = p.acs.nb_aicraft
        sep = min([np.hypot(p.sol_x[a] - p.sol_x[b], p.sol_y[a] - p.sol_y[b]).min() for a in range(n) for b in range(a)] or [np.inf])
        print(f"{nm}[{case}] n_ac={n} N={p.num_nodes}: {dt:.2f}s ticks {info['ticks']} feasible {(info['c_max'] < 1e-5).sum()}/8 cost {p.prob.obj(p.solution):.5e} "
              f"|con| {res:.1e} bounds_ok {info['state_bounds_ok']} min separation {sep:.2f}")
