#!/usr/bin/env python3
"""Latency of ONE C4 problem (16 aircraft x 500 nodes, all-pairs collision) sharded by aircraft over the ranks of a
torchrun job: pack positions -> NCCL all-gather (128 kB) -> shard kernel -> NCCL all-reduce of the cost.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_sharded_c4.py"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "drone-sim-python_b200")]
json_out = os.fdopen(os.dup(1), "w"); os.dup2(2, 1)
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from d2d_b200 import get_engine  # noqa: E402
from d2d_b200.collocation import CollocationProblem, CostSpec  # noqa: E402
from d2d_b200.distributed import ShardedCollocation  # noqa: E402

eng = get_engine()
n_ac, N, h = 16, 500, 0.02
rng = np.random.default_rng(12345)
free = rng.normal(0, 30., 5 * n_ac * N); free[4 * n_ac * N:] = 12. + rng.normal(0, 1, n_ac * N)
cs = lambda: CostSpec(vsp=12., kvel=70., kbank=1., kcol=10., rcol=10., all_pairs=True)
sc = ShardedCollocation(n_ac, N, h, (0., 0.), [], cs(), engine=eng)
fl = eng.to_device(free[sc.shard.idx_free])


def timed(fn, reps=200):
    for _ in range(20):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); e1.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=eng.device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


ms = timed(lambda: sc.evaluate(fl))
res, jac, cost, grad = sc.evaluate(fl)
line = {"what": "one C4 problem sharded by aircraft", "n_gpus": world, "aircraft_per_gpu": n_ac // world, "ms_per_eval": ms, "evals_per_s": 1e3 / ms,
        "cost": float(cost.cpu()[0])}
if rank == 0:
    full = CollocationProblem(n_ac, N, h, cost=cs(), engine=eng)
    fd = eng.to_device(free[None].copy())
    bufs = full.buffers(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(20):
        full.evaluate_device(fd, out=bufs)
    e0.record()
    for _ in range(200):
        full.evaluate_device(fd, out=bufs)
    e1.record(); e1.synchronize()
    line["single_gpu_ms_per_eval"] = e0.elapsed_time(e1) / 200
    line["single_gpu_cost"] = float(bufs["cost"].cpu()[0])
    json_out.write(json.dumps(line) + "\n"); json_out.flush()
dist.barrier()
dist.destroy_process_group()
