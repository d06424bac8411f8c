"""Single-aircraft planner helpers -- the `d2d.opty_utils` call surface (d2d/opty_utils.py:8-187): time grid,
wind, the cost classes IPOPT calls back into (`cost(free, _p)` / `cost_grad(free, _p)`) and the triangle initial
guess.  The cost arithmetic runs in the collocation kernel; `_p` is anything exposing the planner attributes the
reference classes read (`num_nodes`, `obj_scale`, `_slice_*`)."""
import numpy as np

from .collocation import CollocationProblem, CostSpec


def planner_timing(t0, t1, hz, verbose=False):
    """num_nodes, time_step, duration (d2d/opty_utils.py:8-14; note int() truncation of duration*hz)."""
    num_nodes = int((t1 - t0) * hz) + 1
    time_step = 1. / hz
    duration = (num_nodes - 1) * time_step
    if verbose:
        print(f"time_step: {time_step:.3f}s ({hz:.1f}hz), duration {duration:.1f}s -> {num_nodes} nodes")
    return num_nodes, time_step, duration


class WindField:                                             # d2d/opty_utils.py:18-27
    def __init__(self, w=[0., 0.]):
        self.w = w

    def sample_sym(self, _t, _x, _y): return self.w
    def sample_num(self, _t, _x, _y): return self.w
    def __str__(self): return f"{self.w} m/s"


class Aircraft:
    """Names of the state / input trajectories of one aircraft (d2d/opty_utils.py:31-36).  The reference builds
    sympy equations of motion for opty's code generator (`get_eom`, :38-50); here the same equations are the
    collocation kernel, so the class only carries the naming used to order the free vector."""

    def __init__(self, st=None, id=""):
        self.id = id
        self._state_symbols = (f"x{id}(t)", f"y{id}(t)", f"psi{id}(t)")
        self._input_symbols = (f"v{id}", f"phi{id}")

    def get_eom(self, atm, g=9.81):
        """The implicit EoM as text (documentation; the arithmetic lives in csrc/d2dx_colloc.cu)."""
        wx, wy = atm.sample_sym(None, None, None)
        i = self.id
        return (f"x{i}' - v{i} cos(psi{i}) + {wx}", f"y{i}' - v{i} sin(psi{i}) + {wy}", f"psi{i}' - {g}/v{i} tan(phi{i})")


class _EngineCost:
    """Base of the cost classes: subclasses describe themselves as a CostSpec; evaluation is one kernel launch.
    `multi` selects the normalisation of the multi-aircraft classes (d2d/multiopty_utils.py)."""
    multi = False

    def spec(self):
        raise NotImplementedError

    def _problem(self, _p):
        n_ac = _p.acs.nb_aicraft if (self.multi and hasattr(_p, "acs")) else 1
        key = (int(_p.num_nodes), n_ac, float(_p.obj_scale))
        cache = self.__dict__.setdefault("_problems", {})
        if key not in cache:
            cache[key] = CollocationProblem(n_ac, int(_p.num_nodes), 1., cost=self.spec(), obj_scale=float(_p.obj_scale), multi=self.multi)
        return cache[key]

    def cost(self, free, _p):
        return self._problem(_p).obj(free)

    def cost_grad(self, free, _p):
        return self._problem(_p).obj_grad(free)


class CostAirVel(_EngineCost):                               # d2d/opty_utils.py:55-66
    def __init__(self, vsp=10.):
        self.vsp = vsp

    def spec(self): return CostSpec(vsp=self.vsp, kvel=1.)


class CostBank(_EngineCost):                                 # :68-82: mean squared bank (default) or max squared bank
    use_mean = True

    def spec(self): return CostSpec(kbank=1.)

    def _max_mode(self, free, _p, want_cost, want_grad):
        from .engine import get_engine
        eng = get_engine()
        free = np.asarray(free, dtype=np.float64)
        sl = _p._slice_phi
        fd = eng.to_device(np.ascontiguousarray(free.reshape(1, -1)))
        return eng.cost_bank_max(fd, sl.start, sl.stop - sl.start, float(_p.obj_scale), want_cost, want_grad)

    def cost(self, free, _p):
        if self.use_mean:
            return super().cost(free, _p)
        return float(self._max_mode(free, _p, True, False)[0].cpu()[0])          # obj_scale * max(phi^2), :73

    def cost_grad(self, free, _p):
        if self.use_mean:
            return super().cost_grad(free, _p)
        return self._max_mode(free, _p, False, True)[1].cpu().numpy()[0]         # one entry at np.argmax(phi^2), :78-81


class CostInput(_EngineCost):                                # :85-97
    def __init__(self, vsp=10., kvel=1., kbank=1.):
        self.vsp, self.kv, self.kphi = vsp, kvel, kbank

    def spec(self): return CostSpec(vsp=self.vsp, kvel=self.kv, kbank=self.kphi)


class CostObstacle(_EngineCost):                             # :99-134
    def __init__(self, c=(30, 0), r=15., kind=0):
        self.c, self.r, self.kind, self.k = c, r, kind, 2.

    def spec(self): return CostSpec(kobs=1., obstacles=[(self.c[0], self.c[1], self.r)], obs_kind=self.kind)


class CostObstacles(_EngineCost):                            # :136-144
    def __init__(self, obss, kind=0):
        self.obss = [CostObstacle(c=(_o[0], _o[1]), r=_o[2], kind=kind) for _o in obss]
        self.kind = kind

    def spec(self):
        return CostSpec(kobs=1., obstacles=[(o.c[0], o.c[1], o.r) for o in self.obss], obs_kind=self.kind)


class CostComposit(_EngineCost):                             # :147-165
    def __init__(self, obss, vsp=10., kobs=1., kvel=1., kbank=1., obs_kind=0):
        self.kobs, self.kvel, self.kbank, self.vsp, self.obs_kind = kobs, kvel, kbank, vsp, obs_kind
        self.obss = list(obss) if obss is not None else None          # None: "no obstacles" branch (:158-159)

    def spec(self):
        if not self.obss:
            return CostSpec(vsp=self.vsp, kvel=self.kvel, kbank=self.kbank)
        return CostSpec(vsp=self.vsp, kvel=self.kvel, kbank=self.kbank, kobs=self.kobs,
                        obstacles=[(o[0], o[1], o[2]) for o in self.obss], obs_kind=self.obs_kind)


def triangle(p0, p1, va, duration, num_nodes, go_left=1.):
    """Isosceles-triangle initial guess x, y, psi, phi, v over the nodes (d2d/opty_utils.py:171-187).  Host-side
    set-up of an optimiser start point, not part of the evaluated path."""
    p0, p1 = np.asarray(p0, dtype=float), np.asarray(p1, dtype=float)
    leg = p1 - p0
    d = np.linalg.norm(leg)
    u = leg / d
    normal = np.array([-u[1], u[0]])
    D = va * duration
    apex = p0 + leg / 2
    if D > d:
        apex += np.sign(go_left) * np.sqrt(D ** 2 - d ** 2) / 2 * normal
    n1 = int(num_nodes / 2)
    n2 = num_nodes - n1
    pts = np.vstack((np.linspace(p0, apex, n1), np.linspace(apex, p1, n2)))
    out, back = apex - p0, p1 - apex
    psis = np.hstack((np.arctan2(out[1], out[0]) * np.ones(n1), np.arctan2(back[1], back[0]) * np.ones(n2)))
    return pts[:, 0], pts[:, 1], psis, np.zeros(num_nodes), va * np.ones(num_nodes)
