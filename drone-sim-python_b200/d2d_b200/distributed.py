"""Multi-GPU plumbing (one process per GPU, torch.distributed): SURVEY section 8(e).

* Rollouts and batched collocation shard by independent unit (scenario / formation / problem): contiguous
  ranges per rank, NO data-path collective; `reduce_population_stats` is the one final all-reduce.
* One multi-aircraft collocation problem shards by aircraft: the only exchange is an all-gather of the
  aircraft positions (n_ac*N*2 doubles) before the pairwise collision terms and an all-reduce of the scalar
  cost; residual / Jacobian / gradient rows stay with the rank that owns the aircraft.

The compute backend is the engine (CUDA); the class takes it as an argument so the host-side sharding logic can
be exercised on CPU under gloo with a stand-in (tests/test_distributed_cpu.py)."""
import numpy as np
import torch
import torch.distributed as dist

from . import _lib


def shard_range(n_units, world, rank):
    """Contiguous, balanced partition of n_units independent units: [lo, hi) of this rank."""
    base, rem = divmod(int(n_units), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def reduce_population_stats(pop_stats, group=None):
    """pop_stats = [sum of per-scenario sum_sq_err, max of per-scenario max_err] -> whole-job values."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(pop_stats[0:1], op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(pop_stats[1:2], op=dist.ReduceOp.MAX, group=group)
    return pop_stats


class AircraftShard:
    """Index bookkeeping of one rank's slice of an n_ac-aircraft problem (planner layout, 07_multioptyplan.py:41-47)."""

    def __init__(self, n_ac, N, inst, world, rank):
        self.n_ac, self.N, self.world, self.rank = n_ac, N, world, rank
        self.a_lo, self.a_hi = shard_range(n_ac, world, rank)
        self.n_own = self.a_hi - self.a_lo
        own = range(self.a_lo, self.a_hi)
        blk = lambda b: np.arange(b * N, (b + 1) * N)
        self.idx_free = np.concatenate([blk(3 * a + k) for a in own for k in range(3)] +
                                       [blk(3 * n_ac + a) for a in own] + [blk(4 * n_ac + a) for a in own]) if self.n_own else np.zeros(0, np.int64)
        rblk = lambda e: np.arange(e * (N - 1), (e + 1) * (N - 1))
        n_def = 3 * n_ac * (N - 1)
        self.inst_global = [k for k, (var, _, _) in enumerate(inst) if self.a_lo <= var // 3 < self.a_hi]
        self.inst_local = [(inst[k][0] - 3 * self.a_lo, inst[k][1], inst[k][2]) for k in self.inst_global]
        self.idx_con = np.concatenate([rblk(3 * a + e) for a in own for e in range(3)] + [n_def + np.array(self.inst_global, dtype=np.int64)]) \
            if self.n_own else np.zeros(0, np.int64)
        nnz_def = 12 * n_ac * (N - 1)
        self.idx_jac = np.concatenate([np.arange(12 * a * (N - 1), 12 * (a + 1) * (N - 1)) for a in own] +
                                      [nnz_def + np.array(self.inst_global, dtype=np.int64)]) if self.n_own else np.zeros(0, np.int64)


class ShardedCollocation:
    """Aircraft-sharded evaluation of ONE problem across the ranks of `group`.  Every rank must own the same
    number of aircraft (all_gather_into_tensor needs equal chunks)."""

    def __init__(self, n_ac, N, h, wind, inst, cost, obj_scale=1., engine=None, group=None, world=None, rank=None,
                 problem_factory=None):
        self.group = group
        self.world = world if world is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
        self.rank = rank if rank is not None else (dist.get_rank(group) if dist.is_initialized() else 0)
        if n_ac % self.world:
            raise ValueError(f"{n_ac} aircraft do not split evenly over {self.world} ranks")
        self.n_ac, self.N = n_ac, N
        self.shard = AircraftShard(n_ac, N, list(inst), self.world, self.rank)
        if problem_factory is None:
            from .collocation import CollocationProblem
            from .engine import get_engine
            engine = engine or get_engine()
            problem_factory = CollocationProblem
        self.eng = engine
        cost.in_div = n_ac                                   # CostInput normalises by the TOTAL aircraft count
        self.local = problem_factory(self.shard.n_own, N, h, wind=wind, inst=self.shard.inst_local, cost=cost,
                                     obj_scale=obj_scale, layout="compact", multi=True, engine=engine)
        e = self.eng
        self.pos_local = e.empty(self.shard.n_own, 2, N)
        self.pos_all = e.empty(n_ac, 2, N)
        self.res = e.empty(self.local.num_constraints); self.jac = e.empty(self.local.nnz)
        self.grad = e.empty(self.local.num_free); self.cost = e.zeros(1)
        self.scratch = e.colloc_scratch(self.local.c, 1)

    def evaluate(self, free_local, what=_lib.EVAL_ALL):
        """free_local: this rank's slice (device tensor, shard-local planner layout).  Returns
        (residual_local, jac_local, cost_total, grad_local) as device tensors."""
        e = self.eng
        e.colloc_pack_positions(self.shard.n_own, self.N, free_local, self.pos_local)
        if self.world > 1:
            dist.all_gather_into_tensor(self.pos_all, self.pos_local, group=self.group)      # the one exchange step
        else:
            self.pos_all.copy_(self.pos_local)
        e.colloc_eval_shard(self.local.c, self.n_ac, self.shard.a_lo, free_local, self.pos_all, what,
                            self.res, self.jac, self.cost, self.grad, self.scratch)
        if self.world > 1 and (what & _lib.EVAL_COST):
            dist.all_reduce(self.cost, op=dist.ReduceOp.SUM, group=self.group)
        return self.res, self.jac, self.cost, self.grad


def emulate_sharded_eval(n_ac, N, h, wind, inst, cost, free, world, engine=None):
    """Runs every rank's shard kernel in turn on ONE GPU (positions "gathered" by slicing the global free vector)
    and reassembles the global vectors.  Used to check the shard kernel without `world` GPUs."""
    from .collocation import CollocationProblem
    from .engine import get_engine
    import copy
    eng = engine or get_engine()
    inst = list(inst)
    free = np.asarray(free, dtype=np.float64)
    n_con, nnz = 3 * n_ac * (N - 1) + len(inst), 12 * n_ac * (N - 1) + len(inst)
    out = {"residual": np.zeros(n_con), "jac": np.zeros(nnz), "grad": np.zeros(free.size), "cost": 0.}
    pos_all = eng.empty(n_ac, 2, N)
    eng.colloc_pack_positions(n_ac, N, eng.to_device(free), pos_all)
    for rank in range(world):
        sh = AircraftShard(n_ac, N, inst, world, rank)
        c = copy.copy(cost); c.in_div = n_ac
        loc = CollocationProblem(sh.n_own, N, h, wind=wind, inst=sh.inst_local, cost=c, layout="compact", multi=True, engine=eng)
        fl = eng.to_device(free[sh.idx_free])
        res, jac, grad, cst = eng.empty(loc.num_constraints), eng.empty(loc.nnz), eng.empty(loc.num_free), eng.zeros(1)
        eng.colloc_eval_shard(loc.c, n_ac, sh.a_lo, fl, pos_all, _lib.EVAL_ALL, res, jac, cst, grad, eng.colloc_scratch(loc.c, 1))
        out["residual"][sh.idx_con] = res.cpu().numpy()
        out["jac"][sh.idx_jac] = jac.cpu().numpy()
        out["grad"][sh.idx_free] = grad.cpu().numpy()
        out["cost"] += float(cst.cpu().numpy()[0])
    return out
