"""Direct-collocation evaluator on the engine: the callbacks IPOPT drives through
`opty.direct_collocation.Problem` in the reference (06_optyplan.py:62-71, 07_multioptyplan.py:69-78) --
constraints, jacobian, jacobianstructure, objective, gradient -- for one problem or a batch of problems
sharing one description."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .engine import get_engine


from .collocation_spec import CostSpec  # noqa: F401,E402


class CollocationProblem:
    """n_ac aircraft, N nodes, interval h, constant wind, instance constraints [(var, node, value)], cost spec.

    Mirrors the attributes of opty's Problem the reference touches: `num_free`, `num_constraints`,
    `obj(free)`, `obj_grad(free)`, `con(free)`, `con_jac(free)`, `jacobianstructure()`; the IPOPT-style aliases
    `objective/gradient/constraints/jacobian` are provided too.  `free` is a host array (num_free,) or
    (n_prob, num_free); results are fresh host arrays.  `evaluate_device` works on device tensors."""

    def __init__(self, n_ac, N, h, wind=(0., 0.), inst=(), cost=None, obj_scale=1., layout="compact",
                 input_order="numeric", multi=None, engine=None):
        self.eng = engine or get_engine()
        self.n_ac, self.N, self.h = int(n_ac), int(N), float(h)
        self.layout = {"compact": _lib.JAC_COMPACT, "dense": _lib.JAC_OPTY_DENSE, "opty": _lib.JAC_OPTY_DENSE}[layout]
        cost = cost or CostSpec()
        multi = (self.n_ac > 1) if multi is None else multi
        p = _lib.CollocProblem()
        p.n_ac, p.N, p.h = self.n_ac, self.N, self.h
        p.wind[0], p.wind[1] = float(wind[0]), float(wind[1])
        self._keep = []
        if input_order == "opty":                            # opty sorts input trajectories by NAME
            names = [f"phi{i}" for i in range(self.n_ac)] + [f"v{i}" for i in range(self.n_ac)]
            rank = {nm: k for k, nm in enumerate(sorted(names))}
            pphi = np.array([rank[f"phi{i}"] for i in range(self.n_ac)], np.int32)
            pv = np.array([rank[f"v{i}"] for i in range(self.n_ac)], np.int32)
            self._keep += [self.eng.to_device(pphi), self.eng.to_device(pv)]
            p.perm_phi, p.perm_v = self._keep[-2].data_ptr(), self._keep[-1].data_ptr()
        inst = list(inst)
        p.n_inst = len(inst)
        if inst:
            iv = self.eng.to_device(np.array([int(k) for k, _, _ in inst], np.int32))
            inn = self.eng.to_device(np.array([int(n_) for _, n_, _ in inst], np.int32))
            ival = self.eng.to_device(np.array([float(v) for _, _, v in inst], np.float64))
            self._keep += [iv, inn, ival]
            p.inst_var, p.inst_node, p.inst_val = iv.data_ptr(), inn.data_ptr(), ival.data_ptr()
        p.obj_scale, p.vsp, p.kvel, p.kbank = float(obj_scale), cost.vsp, cost.kvel, cost.kbank
        p.in_div = int(cost.in_div) if cost.in_div else (self.n_ac if multi else 1)
        p.kobs, p.obs_kind, p.n_obs = cost.kobs, cost.obs_kind, len(cost.obstacles)
        for k, o in enumerate(cost.obstacles):
            p.obs[k][0], p.obs[k][1], p.obs[k][2] = o
        p.kcol, p.rcol, p.kcol_k = cost.kcol, cost.rcol, cost.kcol_k
        p.col_all_pairs, p.exact_grad = int(cost.all_pairs), int(cost.exact_grad)
        self.c = p
        self.num_free, self.num_constraints, self.nnz = self.eng.colloc_sizes(p, self.layout)
        self._buf = {}

    # ---- device level ---------------------------------------------------------------------------
    def buffers(self, n_prob):
        b = self._buf.get(n_prob)
        if b is None:
            e = self.eng
            b = {"res": e.empty(n_prob, self.num_constraints), "jac": e.empty(n_prob, self.nnz),
                 "cost": e.empty(n_prob), "grad": e.empty(n_prob, self.num_free), "scratch": e.colloc_scratch(self.c, n_prob)}
            if self.layout == _lib.JAC_OPTY_DENSE and self.n_ac > 1:
                e.colloc_init_dense(self.c, n_prob, b["jac"])
            self._buf[n_prob] = b
        return b

    def evaluate_device(self, free_dev, what=_lib.EVAL_ALL, out=None):
        """free_dev: device tensor (n_prob, num_free).  Returns the dict of device output tensors (reused
        between calls unless `out` is given)."""
        n_prob = free_dev.shape[0]
        b = out or self.buffers(n_prob)
        self.eng.colloc_eval(self.c, n_prob, free_dev, self.layout, what, b["res"], b["jac"], b["cost"], b["grad"], b["scratch"])
        return b

    def graph(self, free_dev, what=_lib.EVAL_ALL):
        """Captures one evaluation of `free_dev` (device tensor (n_prob, num_free), updated in place by the caller)
        into a CUDA graph and returns (replay, outputs): `replay()` re-launches it with ~one driver call -- the
        launch-latency-bound single-problem case of an NLP solver's callback loop."""
        n_prob = free_dev.shape[0]
        out = self.buffers(n_prob)
        self.evaluate_device(free_dev, what, out)               # warm up (lazy module load) outside the capture
        torch.cuda.synchronize(self.eng.device)
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(device=self.eng.device)
        with torch.cuda.graph(g, stream=side):
            self.evaluate_device(free_dev, what, out)
        return g.replay, out

    # ---- host level (what IPOPT calls) ----------------------------------------------------------------
    def _run(self, free, what, key):
        free = np.asarray(free, dtype=np.float64)
        single = free.ndim == 1
        fd = self.eng.to_device(np.ascontiguousarray(free.reshape(-1, self.num_free)))
        out = self.evaluate_device(fd, what)[key].cpu().numpy()
        return out[0].copy() if single else out.copy()

    def obj(self, free):
        r = self._run(free, _lib.EVAL_COST, "cost")
        return float(r) if np.ndim(r) == 0 else r

    def obj_grad(self, free): return self._run(free, _lib.EVAL_GRAD, "grad")
    def con(self, free): return self._run(free, _lib.EVAL_RESIDUAL, "res")
    def con_jac(self, free): return self._run(free, _lib.EVAL_JAC, "jac")
    objective, gradient, constraints, jacobian = obj, obj_grad, con, con_jac

    def evaluate(self, free):
        """residual, jacobian values, cost, gradient in one fused launch."""
        free = np.asarray(free, dtype=np.float64)
        single = free.ndim == 1
        fd = self.eng.to_device(np.ascontiguousarray(free.reshape(-1, self.num_free)))
        b = self.evaluate_device(fd, _lib.EVAL_ALL)
        r = tuple(b[k].cpu().numpy().copy() for k in ("res", "jac", "cost", "grad"))
        return tuple(a[0] for a in r) if single else r

    def jacobianstructure(self):
        rows, cols = self.eng.colloc_structure(self.c, self.layout)
        return rows.cpu().numpy(), cols.cpu().numpy()
