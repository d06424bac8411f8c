"""Multi-aircraft planner helpers -- the `d2d.multiopty_utils` call surface (d2d/multiopty_utils.py:9-174)."""
import numpy as np

from . import opty_utils as d2ou
from .collocation import CostSpec
from .opty_utils import _EngineCost


class AircraftSet:                                           # d2d/multiopty_utils.py:9-26
    def __init__(self, n=2):
        self.nb_aicraft = n
        self.aircraft = [d2ou.Aircraft(None, i) for i in range(n)]
        self._state_symbols = tuple(s for ac in self.aircraft for s in ac._state_symbols)
        self._input_symbols = tuple(s for ac in self.aircraft for s in ac._input_symbols)

    def get_eom(self, wind, g=9.81):
        return tuple(e for ac in self.aircraft for e in ac.get_eom(wind, g))


class _MultiCost(_EngineCost):
    multi = True


class CostNull(_MultiCost):                                  # :29-31
    def spec(self): return CostSpec()


class CostAirvel(_MultiCost):                                # :33-43
    def __init__(self, vsp=10.):
        self.vsp = vsp

    def spec(self): return CostSpec(vsp=self.vsp, kvel=1.)


class CostBank(_MultiCost):                                  # :45-53
    def spec(self): return CostSpec(kbank=1.)


class CostInput(_MultiCost):                                 # :55-71
    def __init__(self, vsp=10., kv=1., kphi=1.):
        self.vsp, self.kv, self.kphi = vsp, kv, kphi

    def spec(self): return CostSpec(vsp=self.vsp, kvel=self.kv, kbank=self.kphi)


class CostObstacle(_MultiCost):                              # :74-106 (acts on aircraft 0 only, as upstream)
    def __init__(self, c=(0, 0), r=10., kind=0):
        self.c, self.r, self.kind, self.k = c, r, kind, 2

    def spec(self): return CostSpec(kobs=1., obstacles=[(self.c[0], self.c[1], self.r)], obs_kind=self.kind)


class CostObstacles(_MultiCost):                             # :108-116
    def __init__(self, obss, kind=0):
        self.obss, self.kind = [CostObstacle(c=(_o[0], _o[1]), r=_o[2], kind=kind) for _o in obss], kind

    def spec(self):
        return CostSpec(kobs=1., obstacles=[(o.c[0], o.c[1], o.r) for o in self.obss], obs_kind=self.kind)


class CostCollision(_MultiCost):
    """exp(-(k/r)^2 |p_a - p_b|^2) between aircraft 0 and 1 as upstream (:120-153); `all_pairs=True` extends it to
    every pair, `exact_grad=True` restores the (k/r)^2 factor the upstream gradient omits."""

    def __init__(self, r=3., k=2., all_pairs=False, exact_grad=False):
        self.r, self.k, self.all_pairs, self.exact_grad = r, k, all_pairs, exact_grad

    def spec(self):
        return CostSpec(kcol=1., rcol=self.r, kcol_k=self.k, all_pairs=self.all_pairs, exact_grad=self.exact_grad)


class CostComposit(_MultiCost):                              # :156-174 (NaN weight disables a term, :166-173)
    def __init__(self, kvel=1., kbank=1., kobs=float("nan"), kcol=float("nan"), vsp=10., obss=[], obs_kind=0, rcol=3.,
                 all_pairs=False, exact_grad=False):
        self.kvel, self.kbank, self.kobs, self.kcol = kvel, kbank, kobs, kcol
        self.vsp, self.obss, self.obs_kind, self.rcol = vsp, list(obss), obs_kind, rcol
        self.all_pairs, self.exact_grad = all_pairs, exact_grad

    def spec(self):
        return CostSpec(vsp=self.vsp, kvel=self.kvel, kbank=self.kbank, kobs=self.kobs,
                        obstacles=[(o[0], o[1], o[2]) for o in self.obss], obs_kind=self.obs_kind, kcol=self.kcol,
                        rcol=self.rcol, all_pairs=self.all_pairs, exact_grad=self.exact_grad)
