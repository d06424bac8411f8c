"""Experiments of the multi-aircraft planner -- the classes defined inside 07_multioptyplan.py:170-435 (and re-used by
multi_opt_planner.py:170-242): boundary states per aircraft, grid, cost, bounds, cases.  Same names and values, generated
from a table; `set_case` / `label` behave as upstream (they mutate class attributes)."""
import numpy as np

from . import multiopty_utils as d2mou
from . import opty_utils as d2ou
from .optyplan_scenarios import _derive

deg, nan = np.deg2rad, float("nan")


class exp_0:                                                 # 07_multioptyplan.py:170-202, single aircraft
    name, desc = "exp_0", "single aircraft"
    t0, t1, hz = 0., 10., 50.
    dx, dy = 0., 50.
    p0s = ((0., 0., 0., 0., 10.),)
    p1s = ((dx, dy, np.pi / 2, 0., 10.),)
    wind = d2ou.WindField()
    initial_guess = "tri"
    tol, max_iter = 1e-5, 5000
    vref = 12.
    cost, obj_scale = d2mou.CostInput(vsp=vref, kv=5., kphi=1.), 1.e-1
    x_constraint, y_constraint = (-5, 50), (-5, 50)
    phi_constraint = (-deg(40.), deg(40.))
    v_constraint = (9., 15.)
    obstacles = []
    ncases = 1
    set_case = staticmethod(lambda idx: None)
    label = staticmethod(lambda idx: "")


_Ks = [[1., 1.], [1., 20.], [1., 40.], [1., 60.]]


def _set_K(idx):
    exp_0_1.K = _Ks[idx]
    exp_0_1.cost = d2mou.CostInput(vsp=13., kv=_Ks[idx][0], kphi=_Ks[idx][1])


exp_0_1 = _derive("exp_0_1", exp_0, name="exp_0_1", desc="single aircraft, varying weights", t0=0., p0s=((0., 0., 0., 0., 10.),), t1=10.,
                  p1s=((100, 0, 0., 0., 10.),), x_constraint=None, y_constraint=None, Ks=_Ks, ncases=len(_Ks), set_case=_set_K,
                  label=lambda idx: f"kvel, kbank {exp_0_1.K}")                                                       # :204-217

_dpsi = 0.01
exp_1 = _derive("exp_1", exp_0, name="exp_1", desc="2 aicraft face to face", t1=4.5, vref=12., dpsi=_dpsi,
                p0s=((0., 0., 0., 0., 12.), (50., 0., np.pi - _dpsi, 0., 12.)), p1s=((50., 0., 0., 0., 12.), (0., 0., np.pi + _dpsi, 0., 12.)),
                cost=d2mou.CostInput(vsp=12., kv=5., kphi=1.), obj_scale=1.e-1, x_constraint=None, y_constraint=None, obstacles=[],
                initial_guess="rnd")                                                                                  # :219-233
exp_1_0 = _derive("exp_1_0", exp_1, name="exp_1_0", desc="2 aicraft meeting", t1=4.5, vref=12.,
                  p0s=((0., -20., np.pi / 2, 0., 12.), (7.5, -20., np.pi / 2, 0., 12.)), p1s=((40., 5., 0., 0., 12.), (40., 10., 0, 0., 12.)),
                  cost=d2mou.CostInput(vsp=12., kv=1., kphi=1.), obj_scale=1.e-1, x_constraint=None, y_constraint=None)   # :235-243
exp_1_1 = _derive("exp_1_1", exp_1, name="exp_1_1", desc="2 aicraft face to face, wind", initial_guess="tri")             # :245-250

_d = 5.5 * 12. / 2 / 1.5
exp_2 = _derive("exp_2", exp_0, name="exp_2", desc="4 aicraft", t1=5.5, vref=12., overtime=1.5, d=_d,
                p0s=((-_d, 0., 0., 0., 12.), (_d, 0., np.pi, 0., 12.), (0., _d, -np.pi / 2, 0., 12.), (0., -_d, np.pi / 2, 0., 12.)),
                p1s=((_d, 0., 0., 0., 12.), (-_d, 0., np.pi, 0., 12.), (0., -_d, -np.pi / 2, 0., 12.), (0., _d, np.pi / 2, 0., 12.)),
                cost=d2mou.CostInput(vsp=12., kv=1., kphi=1.), obj_scale=1.)                                          # :252-266

_o3 = ((25, -20, 10),)
exp_3 = _derive("exp_3", exp_0, name="exp_3", desc="single obstacle", t1=6.5, vref=12., p0s=((0., 0., 0., 0., 10.),), p1s=((50., 0., 0., 0., 10.),),
                obstacles=_o3, cx=25, cy=-20, r=10, cost=d2mou.CostObstacle(c=(25, -20), r=10, kind=0), obj_scale=1., x_constraint=None,
                y_constraint=None, v_constraint=(8., 18.), phi_constraint=(-deg(40.), deg(40.)))                        # :268-281
_o31 = ((25, -20, 10), (25, -10, 10))
exp_3_1 = _derive("exp_3_1", exp_3, name="exp_3_1", desc="single obstacle, size/location", obstacles=_o31, ncases=len(_o31),
                  set_case=lambda idx: setattr(exp_3_1, "cost", d2mou.CostObstacle(c=_o31[idx][:2], r=_o31[idx][2], kind=0)),
                  label=lambda idx: f"obstacle {_o31[idx]}")                                                          # :283-291

_o4 = ((30, -10, 20), (70, 15, 20))
exp_4 = _derive("exp_4", exp_0, name="exp_4", desc="set of obstacle", t1=10.5, vref=12., p0s=((0., 0., 0., 0., 10.),), p1s=((100., 0., 0., 0., 10.),),
                obstacles=_o4, cost=d2mou.CostComposit(kvel=1., kbank=1., kobs=1., kcol=nan, vsp=12., obss=_o4, obs_kind=1, rcol=3.), obj_scale=1.,
                x_constraint=None, y_constraint=None, phi_constraint=(-deg(40.), deg(40.)), v_constraint=(9., 18.), initial_guess="rnd")   # :293-309
_o41 = (((30, -10, 15), (30, 25, 15)), ((50, -10, 15), (50, 25, 15)), ((70, -10, 15), (70, 25, 15)))


def _set_o41(idx):
    exp_4_1.obstacles = _o41[idx]
    exp_4_1.cost, exp_4_1.obj_scale = d2mou.CostComposit(kvel=1., kbank=1., kobs=1., kcol=nan, vsp=14., obss=_o41[idx], obs_kind=1, rcol=3.), 1.


exp_4_1 = _derive("exp_4_1", exp_4, name="exp_4_1", desc="set of obstacles, size", v_constraint=(9., 15.), _obstacles=_o41, obj_scale=1e-2,
                  ncases=len(_o41), set_case=_set_o41, label=lambda idx: f"obstacles {_o41[idx]}")                      # :311-325
_dur42 = [9, 10, 11, 12]
exp_4_2 = _derive("exp_4_2", exp_4, name="exp_4_2", desc="set of obstacles, duration", _durations=_dur42, ncases=len(_dur42),
                  set_case=lambda idx: setattr(exp_4_2, "t1", _dur42[idx]), label=lambda idx: f"duration {_dur42[idx]} s",
                  initial_guess="rnd")                                                                                # :327-337


def _set_col(idx):                                           # :355-362
    exp_5.cost, exp_5.obj_scale = d2mou.CostComposit(kvel=70., kbank=1., kobs=nan, kcol=nan if idx == 0 else 10., vsp=exp_5.vref, obss=[],
                                                     obs_kind=0, rcol=3. if idx == 0 else 10.), 1.e0


exp_5 = _derive("exp_5", exp_0, name="exp_5", desc="2 aicraft face to face", t1=4.2, vref=12., dpsi=0.,
                p0s=((0., 0., 0., 0., 12.), (50., 0., np.pi, 0., 12.)), p1s=((50., 0., 0., 0., 12.), (0., 0., np.pi, 0., 12.)),
                x_constraint=None, y_constraint=None, obstacles=[], initial_guess="tri", ncases=2, set_case=_set_col,
                label=lambda idx: f'obj {["Ref", "AntiCol"][idx]}')                                                   # :345-366
exp_5_1 = _derive("exp_5_1", exp_5, name="exp_5", desc="2 aicraft next to one another", t1=8., vref=12.,
                  p0s=((0., 0., 0., 0., 12.), (0., 5., 0, 0., 12.)), p1s=((50., 50., np.pi / 2, 0., 12.), (55., 50., np.pi / 2, 0., 12.)),
                  initial_guess="tri")                                                                                # :368-380

_phi0 = deg(-2.00691223e+01)
_col10 = d2mou.CostComposit(kvel=70., kbank=1., kobs=nan, kcol=10., vsp=12, obss=[], obs_kind=0, rcol=10)
gvf_trial_3ac = _derive("gvf_trial_3ac", exp_5, name="gvf_trial_3ac", desc="Circular formation with 3 aircraft - trial", hz=10, t1=5.5, vref=12, dpsi=0,
                        p0s=((0, 40, 0., _phi0, 12), (25, 20, 0., _phi0, 12), (25, -20, 0., _phi0, 12), (0, -40, 0., _phi0, 12)),
                        p1s=((75, 40, 0, 0, 12), (100, 20, 0, 0, 12), (100, -20, 0, 0, 12), (75, -40, 0, 0, 12)), x_constraint=(-150, 150),
                        y_constraint=(-150, 150), initial_guess="tri", ncases=1, cost=_col10, obj_scale=1.e0)               # :382-403
_t_inf = [10]
inf_traj_4ac = _derive("inf_traj_4ac", exp_5, name="inf trajectory", desc="attempting some fancy inf-like traj", hz=10, t=_t_inf, vref=12, dpsi=0,
                       p0s=((75, 40, 0., deg(20), 12), (100, 40, 0., deg(20), 12), (100, -40, 0., deg(20), 12), (75, -40, 0., deg(20), 12)),
                       p1s=((75, -40, 0., deg(-39), 12), (100, -40, 0., deg(-39), 12), (100, 40, 0., deg(-39), 12), (75, 40, 0., deg(-39), 12)),
                       x_constraint=None, y_constraint=None, initial_guess="tri", ncases=len(_t_inf),
                       set_case=lambda idx: setattr(exp_5, "t1", _t_inf[idx]),       # the cost assignment upstream (:431) is a dead local
                       label=lambda idx: f"t_flight_{_t_inf[idx]}")                                                    # :405-433

from .mission import trap_4  # noqa: E402  (multi_opt_planner.py:221-242)

scens = [exp_0, exp_0_1, exp_1, exp_1_0, exp_1_1, exp_2, exp_3, exp_3_1, exp_4, exp_4_1, exp_4_2, exp_5, exp_5_1, gvf_trial_3ac, inf_traj_4ac]


def desc_all_scens():
    return "\n".join(f"{i}: {s.name} {s.desc}" for i, s in enumerate(scens))


def get_scen(idx):
    return scens[idx]
