#!/usr/bin/env python3
"""Development check of the second-order planner solver on the CPU (host build): every upstream single-vehicle experiment,
multi-start, against the cached IPOPT solutions where the reference ships one (cost of the cached solution recomputed here)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "drone-sim-python_b200"), os.path.join(ROOT, "scripts")]
import ddp_host as d  # noqa: E402
from d2d_b200 import optyplan_scenarios as ops  # noqa: E402
from d2d_b200 import opty_utils as d2ou  # noqa: E402

REF = os.environ.get("D2D_REF", "/root/reference")


def spec_of(exp):
    s = exp.cost.spec()
    return dict(vsp=s.vsp, kvel=s.kvel, kbank=s.kbank, kobs=(s.kobs if s.obstacles else 0.), obstacles=list(s.obstacles), obs_kind=s.obs_kind)


def starts(exp, n):
    lo, hi = exp.phi_constraint
    vs = exp.vref
    out = [(0.0, vs)] + [(f * hi, vs) for f in (0.3, -0.3, 0.7, -0.7, 0.1, -0.1, 0.95, -0.95)]
    rng = np.random.default_rng(0)
    while len(out) < n:
        out.append((rng.uniform(lo, hi), rng.uniform(*exp.v_constraint)))
    return out[:n]


def run(exp, n_starts=9, hz=None, t1=None, opts=None, verbose=True):
    hz = hz or exp.hz
    t1 = t1 if t1 is not None else exp.t1
    N, h, dur = d2ou.planner_timing(exp.t0, t1, hz)
    sp = spec_of(exp)
    box = None
    if exp.x_constraint is not None or exp.y_constraint is not None:
        bx, by = exp.x_constraint or (-1e300, 1e300), exp.y_constraint or (-1e300, 1e300)
        box = (bx[0], bx[1], by[0], by[1], float(os.environ.get("BOXW", "1000")))
    bounds = (exp.phi_constraint[0], exp.phi_constraint[1], exp.v_constraint[0], exp.v_constraint[1])
    w = exp.wind.sample_num(0, 0, 0)
    res = []
    t0 = time.perf_counter()
    for phi0, v0 in starts(exp, n_starts):
        u, xs, info = d.solve(N, h, w, sp["vsp"], sp["kvel"], sp["kbank"], bounds, exp.p0[:3], exp.p1[:3], phi0, v0, obstacles=sp["obstacles"],
                              kobs=sp["kobs"], obs_kind=sp["obs_kind"], box=box, obj_scale=exp.obj_scale, opts=opts)
        inbox = True
        if box is not None:
            inbox = xs[0].min() >= box[0] - 1e-6 and xs[0].max() <= box[1] + 1e-6 and xs[1].min() >= box[2] - 1e-6 and xs[1].max() <= box[3] + 1e-6
        res.append((int(info["flag"]), int(info["iterations"]), info["cost"], info["cmax"], inbox))
    dt = time.perf_counter() - t0
    solved = [r for r in res if r[0] == 2]
    best = min(solved, key=lambda r: r[2]) if solved else None
    if verbose:
        print(f"{exp.__name__:9s} N={N:5d} solved {len(solved)}/{len(res)} its(median solved) {int(np.median([r[1] for r in solved])) if solved else -1:4d} "
              f"best cost {best[2] if best else float('nan'):.6e} in-box {best[4] if best else None} {dt * 1e3:.0f} ms")
    return res, best


def cached_cost(fname, vsp=12.):
    g = np.load(os.path.join(REF, "src", "cache", fname))
    return float(np.mean(np.square(g["sol_v"] - vsp))), len(g["sol_v"])


if __name__ == "__main__":
    OPTS = d.default_options(**{k: (int(v) if k in ("reg_mode", "max_iter", "max_inner", "max_outer", "ls_max") else float(v))
                                for k, v in (a.split("=") for a in sys.argv[1:])})
    _run = run
    run = lambda *a, **kw: _run(*a, opts=OPTS, **kw)
    for e in (ops.exp_0, ops.exp_1, ops.exp_421, ops.exp_2, ops.exp_4, ops.exp_4_1, ops.exp_4_2, ops.exp_5, ops.exp_13, ops.exp_14, ops.exp_0_3, ops.exp_3):
        run(e)
    print("--- against the cached IPOPT solutions (CostAirVel(12), mean (v - 12)^2) ---")
    for fname, exp, t1, hz in (("optyplan_exp0.npz", ops.exp_0, 15., 10.), ("optyplan_exp0_1_0.npz", ops.exp_0, 7., 50.), ("optyplan_exp0_1_1.npz", ops.exp_0, 10., 50.),
                               ("optyplan_exp0_1_2.npz", ops.exp_0, 15., 50.), ("optyplan_exp0_1_3.npz", ops.exp_0, 20., 50.),
                               ("optyplan_exp0_1_4.npz", ops.exp_0, 30., 50.), ("optyplan_exp13 - some traj.npz", ops.exp_13, None, None),
                               ("optyplan_exp 14 - joining 2 points.npz", ops.exp_14, None, None)):
        cc, n = cached_cost(fname)
        res, best = run(exp, t1=t1, hz=hz, verbose=False)
        ok = sum(r[0] == 2 for r in res)
        print(f"{fname:42s} N={n:5d} IPOPT cost {cc:.6e}   ours {best[2] if best else float('nan'):.6e}  solved {ok}/{len(res)}  "
              f"median its {int(np.median([r[1] for r in res if r[0] == 2])) if ok else -1}")
