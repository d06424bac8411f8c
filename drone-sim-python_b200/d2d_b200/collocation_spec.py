"""Cost description shared by the collocation front-ends (no device code)."""
from . import _lib


class CostSpec:
    """Weights of the reference cost classes folded into one description (include/d2dx.h, d2dx_colloc_problem).
    `in_div` = 1 for the single-aircraft classes (d2d/opty_utils.py), n_ac for the multi-aircraft ones
    (d2d/multiopty_utils.py:38,62)."""

    def __init__(self, vsp=10., kvel=0., kbank=0., kobs=float("nan"), obstacles=(), obs_kind=0, kcol=float("nan"),
                 rcol=3., kcol_k=2., all_pairs=False, exact_grad=False, in_div=None):
        self.vsp, self.kvel, self.kbank = float(vsp), float(kvel), float(kbank)
        self.kobs, self.obstacles, self.obs_kind = float(kobs), [tuple(float(v) for v in o) for o in obstacles], int(obs_kind)
        self.kcol, self.rcol, self.kcol_k = float(kcol), float(rcol), float(kcol_k)
        self.all_pairs, self.exact_grad, self.in_div = bool(all_pairs), bool(exact_grad), in_div
        if len(self.obstacles) > _lib.MAX_OBSTACLES:
            raise ValueError(f"at most {_lib.MAX_OBSTACLES} obstacles")
