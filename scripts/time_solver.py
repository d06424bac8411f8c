"""Per-kernel latency of one planner-solve tick at P = 1 and for a population (CUDA events)."""
import sys, numpy as np
sys.path.insert(0, "drone-sim-python_b200")
import torch
from d2d_b200 import _lib, shooting
from d2d_b200.collocation import CollocationProblem, CostSpec
from d2d_b200.engine import get_engine
eng = get_engine()

def timed(fn, reps=200):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps

for (P, n_ac, N, h) in ((1, 1, 101, 0.1), (1, 1, 1001, 0.02), (1, 16, 500, 0.02), (2048, 1, 101, 0.1), (16384, 1, 101, 0.1), (1024, 1, 1001, 0.02), (16384, 1, 1001, 0.02), (512, 16, 500, 0.02)):
    prob = CollocationProblem(n_ac, N, h, cost=CostSpec(vsp=12., kvel=1., kbank=1.), multi=n_ac > 1)
    rng = np.random.default_rng(1)
    p1 = np.stack([rng.uniform(-10, 10, (P, n_ac)), rng.uniform(28, 40, (P, n_ac)), np.pi + rng.uniform(-0.5, 0.5, (P, n_ac))], 1)
    nlp = shooting.ShootingNLP(prob, np.zeros((3, n_ac)), p1, (-0.52, 0.52), (9., 14.), P=P)
    th = nlp.theta_of(np.full((n_ac, N), 0.1), np.full((n_ac, N), 12.))
    o = _lib.LbfgsOptions(m=20, max_inner=500, max_outer=30, ls_max=30, window=10, gtol=1e-10, ftol=1e-10, ctol=1e-8, rho0=10., rho_max=1e6)
    off = eng.lbfgs_layout(P, nlp.n, nlp.n_con, o)
    state, nrun, xt = eng.empty(off[0]), eng.zeros(1, dtype=torch.int32), th.clone()
    eng.lbfgs_init(P, nlp.n, nlp.n_con, o, state, nlp.lam, nlp.rho)
    e = eng
    fwd = lambda: e.shoot_forward(nlp.c_prob, P, xt, nlp.bounds, nlp.p0, nlp.p1, nlp.u_phys, nlp.xs, nlp.c)
    adj = lambda: e.shoot_adjoint(nlp.c_prob, P, xt, nlp.bounds, nlp.u_phys, nlp.xs, nlp.c, nlp.lam, nlp.rho, nlp.cost_ac, nlp.lagr_ac, nlp.grad)
    tick = lambda: e.al_lbfgs_tick(P, nlp.n, nlp.n_con, o, state, xt, nlp.lagr_ac, nlp.cost_ac, n_ac, nlp.grad, nlp.c, nlp.lam, nlp.rho, nrun)
    for _ in range(60):                                   # fill the history so that the tick does a full two-loop
        fwd(); adj(); tick()
    t_f, t_a = timed(fwd), timed(adj)
    def full():
        fwd(); adj(); tick()
    t_all = timed(full)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=torch.cuda.Stream(device=e.device)):
        for _ in range(16): full()
    t_graph = timed(g.replay, 20) / 16
    print(f"P={P} n_ac={n_ac} N={N} n={nlp.n}: forward {t_f:.1f} us  adjoint {t_a:.1f} us  eval+tick {t_all:.1f} us  in graph {t_graph:.1f} us/tick "
          f"-> {P / t_graph * 1e6:.3g} problem-ticks/s")
