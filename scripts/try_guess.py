import sys, time, numpy as np
sys.path.insert(0, "drone-sim-python_b200")
import torch
from d2d_b200 import planner as pl, optyplan_scenarios as S, multiopty_scenarios as M
class exp_c3(pl.exp_0):
    t1, hz = 20., 50.
cases = [("exp_0", pl.Planner, pl.exp_0), ("c3", pl.Planner, exp_c3), ("exp_0_3", pl.Planner, S.exp_0_3), ("exp_3", pl.Planner, S.exp_3), ("exp_4_2", pl.Planner, S.exp_4_2),
         ("exp_14", pl.Planner, S.exp_14), ("m_exp_2", pl.MultiPlanner, M.exp_2), ("gvf3", pl.MultiPlanner, M.gvf_trial_3ac), ("m_exp_5_1", pl.MultiPlanner, M.exp_5_1)]
pl.Planner(pl.exp_0).run()
for nm, cls, exp in cases:
    exp.set_case(min(1, exp.ncases - 1))
    for gp in (False, True):
        p = cls(exp); p.configure(tol=1e-6)
        t0 = time.time(); info = p.run(n_starts=8, guess_from_path=gp); torch.cuda.synchronize(); dt = time.time() - t0
        xs = np.concatenate([np.ravel(s) for s in (p.sol_x if isinstance(p.sol_x, list) else [p.sol_x])]); ys = np.concatenate([np.ravel(s) for s in (p.sol_y if isinstance(p.sol_y, list) else [p.sol_y])])
        print(f"{nm:10s} from_path={gp!s:5s} {dt:5.2f}s ticks {info['ticks']:6d} its[0] {info['iterations_each'][0]:5d} feasible {(info['c_max'] < 1e-5).sum()}/8 start0 {'ok' if info['c_max'][0] < 1e-5 else 'no'} "
              f"cost {p.prob.obj(p.solution):.5e} bounds_ok {info['state_bounds_ok']} x[{xs.min():.0f},{xs.max():.0f}] y[{ys.min():.0f},{ys.max():.0f}]")
