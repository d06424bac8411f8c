"""Runs the single-vehicle planner experiments of optyplan_scenarios through Planner.run and reports feasibility."""
import sys, time, numpy as np
sys.path.insert(0, "drone-sim-python_b200")
import torch
from d2d_b200 import planner as pl, optyplan_scenarios as S, shooting
import os, functools
if os.environ.get("DDP_RETRIES"):                       # A/B of the retry ladder's length
    shooting.solve_ddp = functools.partial(shooting.solve_ddp, retries=int(os.environ["DDP_RETRIES"]))
W = float(sys.argv.pop(1)) if len(sys.argv) > 1 and sys.argv[1][0].isdigit() else 100.
names = sys.argv[1:] or ["exp_0", "exp_0_3", "exp_1", "exp_2", "exp_3", "exp_4", "exp_4_1", "exp_4_2", "exp_5", "exp_13", "exp_14"]
for nm in names:
    exp = getattr(S, nm)
    for case in range(min(exp.ncases, 2)):
        exp.set_case(case)
        p = pl.Planner(exp)
        p.configure(tol=exp.tol, max_iter=exp.max_iter)
        t0 = time.time(); info = p.run(n_starts=8, state_weight=W); torch.cuda.synchronize(); dt = time.time() - t0
        res = np.abs(p.prob.con(p.solution)).max()
        print(f"{nm}[{case}] N={p.num_nodes}: {dt:.2f}s {info.get('method')} its {info.get('iterations')} feasible starts {(info['c_max'] < 1e-5).sum()}/{len(info['c_max'])}  best cost {p.prob.obj(p.solution):.5e} "
              f"|con| {res:.1e} bounds_ok {info['state_bounds_ok']}  x [{p.sol_x.min():.1f},{p.sol_x.max():.1f}] y [{p.sol_y.min():.1f},{p.sol_y.max():.1f}]")
