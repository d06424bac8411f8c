#!/usr/bin/env python3
"""Small single-kernel drivers for ncu captures (round 2): python scripts/prof_target.py <target> [reps]
targets: c4batch (256 x C4 all-pairs), c3batch (4096 x C3), rollout (1e6 circles x 400 steps, log x100), poly, composite, formation, tracker"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "drone-sim-python_b200")]
import torch  # noqa: E402
from d2d_b200 import _lib, get_engine  # noqa: E402
from d2d_b200.collocation import CollocationProblem, CostSpec  # noqa: E402

target = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
eng = get_engine()
rng = np.random.default_rng(12345)


def timed(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / reps


if target in ("c4batch", "c3batch"):
    n_ac, N, h, n_prob, cost = (16, 500, 0.02, 256, CostSpec(vsp=12., kvel=70., kbank=1., kcol=10., rcol=10., all_pairs=True)) if target == "c4batch" \
        else (1, 1001, 0.02, 4096, CostSpec(vsp=12., kvel=1.))
    prob = CollocationProblem(n_ac, N, h, inst=[(k, 0, 0.) for k in range(3 * n_ac)], cost=cost)
    free = eng.to_device(rng.normal(0, 3., (n_prob, prob.num_free)) + 12. * (np.arange(prob.num_free) >= 4 * n_ac * N))
    bufs = prob.buffers(n_prob)
    ms = timed(lambda: prob.evaluate_device(free, _lib.EVAL_ALL, bufs))
    print(f"{target}: {ms * 1e3:.1f} us per launch, {n_ac * N * n_prob} aircraft-nodes")
elif target in ("rollout", "poly"):
    sys.path.insert(0, ROOT)
    import bench
    from d2d_b200.simulation import MonteCarloRollout
    B, T = int(os.environ.get("PROF_B", 10 ** 6)), 400
    if target == "rollout":
        w = bench.workload(B, 12345)
        X0 = bench.flat_state0(w) + w["noise"]
        mc = MonteCarloRollout(B, np.arange(T + 1) * 0.01, _lib.SEG_CIRCLE, log_every=100, n_chunks=1, host_log=False)
        par = np.zeros((6, B)); par[1], par[2], par[3], par[4], par[5] = w["cx"], w["cy"], w["r"], w["v"] / w["r"], w["a0"]
        mc.set_inputs(par, w["wind"], X0)
    else:
        from d2d_b200 import trajectory as ddt
        B = 500000
        Y0 = np.zeros((B, 2, 4)); Y1 = np.zeros((B, 2, 4))
        a0, a1 = rng.uniform(-0.5, 0.5, B), rng.uniform(1.0, 2.0, B)
        Y0[:, 0, 0], Y0[:, 1, 0] = rng.uniform(-20, 20, B), rng.uniform(-20, 20, B)
        Y0[:, 0, 1], Y0[:, 1, 1] = 10 * np.cos(a0), 10 * np.sin(a0)
        Y1[:, 0, 0], Y1[:, 1, 0] = Y0[:, 0, 0] + rng.uniform(150, 250, B), Y0[:, 1, 0] + rng.uniform(150, 250, B)
        Y1[:, 0, 1], Y1[:, 1, 1] = 10 * np.cos(a1), 10 * np.sin(a1)
        msb = ddt.MinSnapBatch.from_boundaries(Y0, Y1, 33.65)
        mc = MonteCarloRollout(B, np.arange(T + 1) * 0.01, _lib.SEG_POLY, log_every=100, n_chunks=1, host_log=False)
        parp = np.zeros((_lib.SEG_NPAR, B)); parp[1:9] = msb.coefs0[:, 0].T; parp[9:17] = msb.coefs0[:, 1].T
        mc.set_inputs(parp, np.zeros((B, 2)), np.stack([Y0[:, 0, 0] + 1., Y0[:, 1, 0] - 1., a0, 0 * a0, 0 * a0 + 10.], 1))
    mc.upload()
    ms = timed(mc.run_device)
    print(f"{target}: {ms:.3f} ms per launch, {B * T} aircraft-steps, {B * T / ms / 1e6:.2f} G steps/s")
elif target == "formation":
    from d2d_b200.simulation import chain_incidence
    n_ac, T = 6, 1200
    F = eng.sm_count * 5 * (eng.formation_threads_per_sm // 32)
    M = F * n_ac
    X0 = eng.to_device(np.ascontiguousarray(np.tile(np.array([20, 30, -np.pi / 2, 0, 10.]), (M, 1)).T))
    c, r, ac = eng.zeros(2, M), eng.to_device(np.full(M, 60.)), eng.to_device(np.stack([np.full(M, 0.01), np.full(M, 1.)]))
    z = np.ones(n_ac - 1) * 2 * np.pi / n_ac
    Xf = eng.empty(5, M)
    ms = timed(lambda: eng.rollout_formation(n_ac, chain_incidence(n_ac), z, X0, c, r, ac, 4e-4, 15, 20, 15., 0.05, 0, T - 1, 5, X_final=Xf))
    print(f"formation: {ms:.3f} ms, {M * (T - 1)} aircraft-steps")
elif target == "tracker":
    Mt, Tt = eng.sm_count * 1024, 101
    tt = np.arange(Tt) * 0.1
    rr, vv = rng.uniform(30, 60, Mt), rng.uniform(10, 14, Mt)
    om = vv / rr
    al = om[None, :] * tt[:, None]
    ref = np.stack([rr * np.cos(al), rr * np.sin(al), -vv * np.sin(al), vv * np.cos(al), -vv * om * np.cos(al), -vv * om * np.sin(al)], 1)
    X0t = eng.to_device(np.ascontiguousarray(np.stack([rr + 1., 0 * rr - 1., 0 * rr + np.pi / 2, 0 * rr, vv], 0)))
    refd, wz, act, Xft = eng.to_device(ref), eng.zeros(2, Mt), eng.to_device(np.stack([np.full(Mt, 0.01), np.full(Mt, 1.)])), eng.empty(5, Mt)
    ms = timed(lambda: eng.rollout_tracker(refd, X0t, wz, act, 0.1, 0, Tt - 1, 10, X_final=Xft))
    print(f"tracker: {ms:.3f} ms, {Mt * (Tt - 1)} aircraft-steps, {Mt * (Tt - 1) / ms / 1e6:.2f} G steps/s")
elif target == "composite":
    from d2d_b200 import scenario as dds, simulation
    scen = dds.get("patrol_3")
    B = 148 * 512
    trajs = [scen.trajs[k % len(scen.trajs)] for k in range(B)]
    X0 = np.stack([np.asarray(scen.X0s[k % len(scen.trajs)], dtype=np.float64) for k in range(B)]) + rng.normal(0, 1, (B, 5)) * np.array([2, 2, .1, .02, .2])
    W = scen.windfield.sample(0, None)
    import time as _t
    t0 = _t.perf_counter()
    res = simulation.rollout(scen.time, trajs, W, X0, return_log=False)
    torch.cuda.synchronize()
    print(f"composite: {B} x {len(scen.time) - 1} steps in {(_t.perf_counter() - t0) * 1e3:.1f} ms wall (incl. packing)")
