#!/usr/bin/env python3
"""Timing of the second-order planner solve on the GPU: C3 (exp_0 on the 20 s / 50 Hz grid) with 9 and 64 starts, a population of
4096 exp_0-grid problems.    python scripts/time_ddp.py"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "drone-sim-python_b200")]
from d2d_b200 import planner as pl  # noqa: E402
from d2d_b200.shooting import ShootingNLP, solve_ddp  # noqa: E402


class exp_c3(pl.exp_0):
    t1, hz = 20., 50.


pl.Planner(pl.exp_0).run(method="ddp")                               # module load
for n in (1, 64):
    p = pl.Planner(exp_c3); p.configure(tol=1e-8)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    info = p.run(n_starts=n, method="ddp")
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"C3 n_starts={n}: {dt:.3f} s, starts {len(info['c_max'])}, feasible {(info['c_max'] < 1e-6).sum()}, median sweeps {np.median(info['iterations_each'])}, "
          f"cost {info['cost'][info['best']]:.3e}")
rng = np.random.default_rng(12345)
Pp = 4096
pe = pl.Planner(pl.exp_0)
p1 = np.stack([rng.uniform(-10, 10, Pp), rng.uniform(28, 40, Pp), np.pi + rng.uniform(-0.5, 0.5, Pp)], 1).reshape(Pp, 3, 1)
nlp = ShootingNLP(pe.prob, np.zeros((3, 1)), p1, pl.exp_0.phi_constraint, pl.exp_0.v_constraint, P=Pp)
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    _, info = solve_ddp(nlp, 0.1, 12., ctol=1e-8)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"population {Pp}: {dt:.3f} s, solved {(info['flag'] == 2).sum()}, median sweeps {np.median(info['iterations_each'])}")
