"""Experiments of the single-vehicle planner -- the configuration source `d2d/optyplan_scenarios.py` (classes with
mutable class attributes: t0, p0, t1, p1, hz, wind, cost, obj_scale, input and state bounds, cases).  Same names and
values, generated from a table; `set_case(idx)` of the multi-case experiments mutates class attributes like upstream
(exp_0_1 / exp_0_2 / exp_6 write into exp_0, which their siblings inherit from)."""
import numpy as np

from . import opty_utils as d2ou
from .planner import exp_0

deg = np.deg2rad
exp_0.set_case = staticmethod(lambda idx: None)               # d2d/optyplan_scenarios.py:27-28
exp_0.label = staticmethod(lambda idx: "")


def _derive(cls_name, base, **attrs):
    """subclass `base` with overridden class attributes; callables become static methods (upstream calls them on the class)."""
    body = {k: (staticmethod(v) if callable(v) and not isinstance(v, type) and k in ("set_case", "label") else v) for k, v in attrs.items()}
    return type(cls_name, (base,), body)


_t1s = [7., 10., 15., 20, 30]
exp_0_1 = _derive("exp_0_1", exp_0, name="exp0_1", desc="changing duration", tol=1e-5, max_iter=5000, t1s=_t1s, ncases=len(_t1s),
                  set_case=lambda idx: setattr(exp_0, "t1", _t1s[idx]), label=lambda idx: f"{_t1s[idx]:.1f} s")            # :30-42

_winds = [[0., 0.], [1., 0.], [2., 0.], [5., 0.]]
exp_0_2 = _derive("exp_0_2", exp_0, name="exp0_2", desc="changing wind", tol=1e-5, max_iter=5000, winds=_winds, ncases=len(_winds),
                  set_case=lambda idx: setattr(exp_0, "wind", d2ou.WindField(w=_winds[idx])),
                  label=lambda idx: f"wind {_winds[idx]} m/s")                                                          # :44-53

exp_0_3 = _derive("exp_0_3", exp_0, name="exp0_3", desc="xy constraints", tol=1e-5, max_iter=5000, cost=d2ou.CostBank(), obj_scale=1.e-1,
                  x_constraint=(-5., 45.), y_constraint=(-1., 51.), t1=20.)                                              # :56-63

exp_1 = _derive("exp_1", exp_0, name="exp_1", desc="combined phi/vel objective", t0=0., p0=(0., 0., 0., 0., 12.), t1=10.,
                p1=(100., 0., 0., 0., 12.), cost=d2ou.CostInput(vsp=12., kvel=1., kbank=50.), obj_scale=1.)               # :66-72

_Ks = [[1., 0.5], [1., 1.], [1., 10.], [1., 20.], [1., 30.], [1., 40.], [1., 50.]]


def _set_K(idx):
    exp_1_1.K = _Ks[idx]
    exp_1_1.cost = d2ou.CostInput(vsp=12., kvel=_Ks[idx][0], kbank=_Ks[idx][1])


exp_1_1 = _derive("exp_1_1", exp_1, name="exp_1_1", desc="combined phi/vel objective", Ks=_Ks, ncases=len(_Ks), set_case=_set_K,
                  label=lambda idx: f"kvel, kbank {exp_1_1.K}")                                                         # :75-84

exp_421 = _derive("exp_421", exp_0, cost=d2ou.CostBank(), name="exp1", desc="min mean bank objective")                   # :86-89
exp_2 = _derive("exp_2", exp_0, cost=d2ou.CostComposit(None, 11., kobs=0., kvel=0.1, kbank=10.), name="exp2", desc="bank/vel obective")   # :91-94
exp_3 = _derive("exp_3", exp_2, cost=d2ou.CostComposit(None, 12., kobs=0., kvel=0.5, kbank=1.), obj_scale=1., x_constraint=(-5., 35.),
                y_constraint=(-5., 35.), t1=20., name="exp3", desc="bank/vel obective, xy constraints")                   # :96-103

_obs4 = ((25, -20, 10),)
exp_4 = _derive("exp_4", exp_0, t0=0., p0=(0., 0., 0, 0., 10.), t1=6.5, p1=(50., 0., 0, 0., 10.), obstacles=_obs4,
                cost=d2ou.CostComposit(_obs4, vsp=15., kobs=0.5, kvel=0.5, kbank=1.), obj_scale=1.e-2, phi_constraint=(-deg(40.), deg(40.)),
                x_constraint=(-5., 105.), y_constraint=(-15., 35.), v_constraint=(9., 15.), name="exp4", desc="obstacle - simple case")   # :105-116

exp_4_1 = _derive("exp_4_1", exp_0, t0=0., p0=(0., 0., 0, 0., 10.), t1=8.5, p1=(100., 0., 0, 0., 10.), obstacles=((50, -10, 25),),
                  cost=d2ou.CostInput(vsp=12., kvel=0.5, kbank=1.), obj_scale=1.e-2, x_constraint=(-5., 105.), y_constraint=(-10., 40.),
                  phi_constraint=(-deg(40.), deg(40.)), v_constraint=(9., 15.), name="exp4_1", desc="obstacle - simple case")      # :118-133

_maze = ((25, 0, 15), (55, 7.5, 12), (80, -10, 12))
_maze_attrs = dict(t1=15., p1=(100., 0., 0, 0., 10.), obstacles=_maze, cost=d2ou.CostComposit(_maze, vsp=15., kobs=0.5, kvel=0.5, kbank=1.),
                   phi_constraint=(-deg(40.), deg(40.)), obj_scale=1.e-2, x_constraint=(-5., 105.), y_constraint=(-15., 35.),
                   v_constraint=(9., 15.), name="exp4", desc="obstacles - maze")
exp_4_2 = _derive("exp_4_2", exp_0, **_maze_attrs)                                                                       # :135-147
exp_4_3 = _derive("exp_4_3", exp_0, **_maze_attrs)                                                                       # :149-161

_checker = [(i * 20., j * 20., 10.) for i in range(5) for j in range(5) if (i + j) % 2]
exp_5 = _derive("exp_5", exp_0, t0=0., p0=(0., 40., 0, 0., 10.), t1=12., p1=(100., 40., 0, 0., 10.), obstacles=_checker,
                cost=d2ou.CostComposit(_checker, vsp=15., kobs=0.5, kvel=10., kbank=1.), phi_constraint=(-deg(40.), deg(40.)), name="exp5")   # :163-178

_p0s6 = ((0, 10, np.pi / 2, 0., 10.), (0, 20, np.pi / 2, 0., 10.), (10, 10, np.pi / 2, 0., 10.), (10, 20, np.pi / 2, 0., 10.),
         (20, 10, np.pi / 2, 0., 10.), (20, 20, np.pi / 2, 0., 10.))
_p1s6 = [(0, 50, np.pi, 0., 10.) for _ in _p0s6]


def _set_rdv(idx):
    exp_0.p0, exp_0.p1 = _p0s6[idx], _p1s6[idx]


def _make_exp_6():
    """Upstream sets exp_0.t1 = 15 while the class body of exp_6 is executed (:205); done when exp_6 is first asked for, so
    that importing this module leaves exp_0 as the planner front ends define it."""
    exp_0.t1 = 15.
    return _derive("exp_6", exp_0, p0s=_p0s6, p1s=_p1s6, ncases=len(_p0s6), x_constraint=(-50., 105.), y_constraint=(0, 100), set_case=_set_rdv,
                   label=lambda idx: f"{idx}", name="exp6", desc="Rendez-vous")                                          # :180-212


class exp_13:                                                # :214-231 (stands alone, not derived from exp_0)
    ncases = 1
    tol, max_iter = 1e-5, 1500
    vref = 12.
    cost = d2ou.CostAirVel(vref)
    obj_scale = 1.
    wind = d2ou.WindField(w=[0., 0.])
    obstacles = ()
    t0, p0 = 0., (75, 40, deg(0), 0, 12)
    t1, p1 = 3., (100, 20, deg(-90), 0, 12)
    x_constraint, y_constraint = None, None
    phi_constraint = (-deg(30.), deg(30.))
    v_constraint = (9., 14.)
    hz = 10.
    name, desc = "exp13 - some traj", "just going"
    set_case = staticmethod(lambda idx: None)
    label = staticmethod(lambda idx: "")


exp_14 = _derive("exp_14", exp_0, name="exp 14 - joining 2 points", desc="single ac traj computation for test case 2 of full sim", ncases=1,
                 tol=1e-5, max_iter=1500, vref=12, cost=d2ou.CostAirVel(12), obj_scale=1, wind=d2ou.WindField(w=[0, 0]), obstacles=(),
                 t0=0, p0=(-49.98, -58.14, 2.22, -0.35, 15.), t1=12, p1=(75, 40, 0, 0, 12), x_constraint=(-150, 150), y_constraint=(-150, 150),
                 v_constraint=(9., 15.), phi_constraint=(-deg(40.), deg(40.)), initial_guess="tri", hz=10)               # :233-253


def __getattr__(name):                                       # exp_6 and the list that contains it are built on first use
    if name == "exp_6":
        globals()["exp_6"] = _make_exp_6()
        return globals()["exp_6"]
    if name == "scens":
        e6 = __getattr__("exp_6") if "exp_6" not in globals() else globals()["exp_6"]
        globals()["scens"] = [exp_0, exp_0_1, exp_0_2, exp_0_3, exp_1, exp_1_1, exp_2, exp_3, exp_4, exp_4_1, exp_4_2, exp_5, e6, exp_13, exp_14]
        return globals()["scens"]
    raise AttributeError(name)


def desc_all():
    return "\n".join(f"{i}: {s.name} {s.desc}" for i, s in enumerate(__getattr__("scens") if "scens" not in globals() else globals()["scens"]))


def desc_one(idx):
    s = (__getattr__("scens") if "scens" not in globals() else globals()["scens"])[idx]
    return f"{s.name} {s.desc}\ninitial state {s.t0} {s.p0}\nfinal state {s.t1} {s.p1}\n"
