"""Named scenarios -- the `d2d.scenario` registry (d2d/scenario.py): trajectories + time grid + wind + initial
states + perturbations.  Data only; `d2d_b200.simulation.test_simulation` runs one on the engine.
"circle" defaults, like upstream, to the constant-air-speed space-indexed circle (TrajSiSpline: its constructor runs the
upstream optimiser with the flat output evaluated on the engine); `cst_gvel=True` gives the plain TrajectoryCircle."""
import numpy as np

from . import trajectory as ddt
from . import trajectory_factory as ddtf
from .dynamic import Aircraft
from .guidance import DiffFlatness, WindField

_scenarios = {}
_default_dt = 0.01                                           # d2d/scenario.py:21


def register(S):
    _scenarios[S.name] = (S.desc, S)
    return S


def list_available():
    return [f"{k}: {v[0]}" for k, v in sorted(_scenarios.items())]


class Scenario:
    """Fills whatever a subclass left unset (d2d/scenario.py:24-52)."""

    def __init__(self):
        nv = len(self.trajs)
        if not hasattr(self, "time"):
            self.time = np.arange(0., np.max([tr.duration for tr in self.trajs]), _default_dt)
        if not hasattr(self, "aircrafts"):
            self.aircrafts = [Aircraft() for _ in range(nv)]
        if not hasattr(self, "perts"):
            self.perts = [np.zeros((len(self.time), Aircraft.s_size)) for _ in range(nv)]
        if not hasattr(self, "windfield"):
            self.windfield = WindField()
        if not hasattr(self, "X0s"):
            t0 = self.time[0]
            Ys = np.stack([tr.get(t0) for tr in self.trajs])
            W = self.windfield.sample(t0, None)
            self.X0s = list(DiffFlatness.state_and_input_from_output(Ys, W, self.aircrafts[0])[0])
        if not hasattr(self, "extends"):
            self.extends = (0., 100., 0., 100.)
        if not hasattr(self, "ppctl"):
            self.ppctl = False

    def summarize(self):
        ext = "".join(f"{e:.1f} " for e in self.extends)
        return (f"{len(self.trajs)} trajectories\nduration: {self.time[-1] - self.time[0]:.2f}s\n"
                f"wind: {self.windfield.summarize()}\nextends: {ext}")


@register
class ScenLine(Scenario):                                    # d2d/scenario.py:72-85
    name = desc = "line"

    def __init__(self):
        self.trajs = [ddt.TrajectoryLine([0, 25], [100, 25], v=10., t0=0.)]
        self.extends = (-10, 110, 0, 50)
        self.windfield = WindField()
        self.time = np.arange(0, 12., 0.01)
        self.X0s = [[10, 10, 0, 0, 10]]
        self.perts = [np.zeros((len(self.time), Aircraft.s_size))]
        self.perts[0][600, Aircraft.s_y] = 10
        super().__init__()


@register
class ScenLine2(Scenario):                                   # :88-98
    name = desc = "line2"

    def __init__(self):
        self.trajs = [ddtf.TrajTwoLines()]
        self.extends = self.trajs[0].extends
        self.windfield = WindField()
        self.time = np.arange(0, 12., 0.01)
        self.X0s = [[0, 10, 0, 0, 10]]
        super().__init__()


@register
class ScenCircle(Scenario):                                  # :101-139
    name = desc = "circle"

    def __init__(self, duration=None, cst_gvel=False, knots=None):
        if cst_gvel:                                         # constant ground speed (:106-107)
            self.trajs = [ddt.TrajectoryCircle(alpha0=3 * np.pi / 2)]
        else:                                                # constant air speed (:108-109): upstream's default
            self.trajs = [ddtf.TrajSiSpline(duration=20., knots=knots)]
        self.extends = (-10, 75, -10, 75)
        self.windfield = WindField([5, 0])
        self.time = np.arange(0, self.trajs[0].duration, 0.01)
        super().__init__()


@register
class ScenSquare(Scenario):                                  # :142-151
    name = desc = "square"

    def __init__(self):
        self.trajs = [ddtf.TrajSquare()]
        self.extends = self.trajs[0].extends
        self.windfield = WindField()
        self.time = np.arange(0, 30., 0.01)
        self.X0s = [[0, 0, 0, 0, 10]]
        super().__init__()


@register
class ScenMultiCircle(Scenario):                             # :155-168
    name, desc = "mucir", "5 circles (30m radius, x offset)"

    def __init__(self, dx=0., dalpha=np.deg2rad(30.), nc=5, v=10.):
        self.trajs = [ddt.TrajectoryCircle(c=[40. + i * dx, 50.], alpha0=i * dalpha, v=v) for i in range(nc)]
        self.extends = (0, 80, 10, 90)
        self.X0s = [[75 - 5 * i, 60 + 5 * i, np.pi, 0, 10] for i in range(nc)]
        self.windfield = WindField([5, 0])
        self.time = np.arange(0, 20, 0.01)
        super().__init__()


@register
class ScenMultiCircle2(Scenario):                            # :172-184
    name, desc = "mucir2", "2 circles (30m radius, x offset)"

    def __init__(self, dx=0):
        self.trajs = [ddt.TrajectoryCircle(c=[40., 50.], alpha0=np.deg2rad(0.)),
                      ddt.TrajectoryCircle(c=[40. + dx, 50.], alpha0=np.deg2rad(30.))]
        self.extends = (0, 100, 0, 100)
        self.X0s = [[75, 50, np.pi / 2, 0, 10], [85, 70, np.pi / 1.5, 0, 10]]
        self.windfield = WindField([1., 0])
        self.time = np.arange(0, 18, 0.01)
        super().__init__()


@register
class ScenPatrol(Scenario):                                  # :189-205
    name, desc = "patrol", "The original 'Line Patrol' scenario"

    def __init__(self):
        self.trajs = [ddtf.TrajLineWithIntro(Y0=[0., 100.], Y1=[0., 50.], Y2=[200., 50.], r=25.),
                      ddtf.TrajLineWithIntro(Y0=[0., 0.], Y1=[0., 50.], Y2=[200., 50.], r=-25.)]
        self.X0s = [[0, 100, -np.pi, 0, 10], [0, 0, -np.pi, 0, 10]]
        self.extends = (-30, 110, -10, 110)
        self.windfield = WindField([0, 2.5])
        self.time = np.arange(0, 20, 0.01)
        super().__init__()


@register
class ScenPatrol2(Scenario):                                 # :209-227
    name, desc = "patrol_2", "dev patrol"

    def __init__(self):
        lead = ddtf.TrajLineWithIntro(Y0=[0., 100.], Y1=[0., 50.], Y2=[100., 50.], r=25.)
        w1 = ddtf.TrajWithIntro([-20, 0], ddtf.TrajSlalom(p1=[0, 50], p2=[100, 50], v=10.), duration=8.)
        w2 = ddtf.TrajWithIntro([0, 0], ddtf.TrajSlalom(p1=[0, 40], p2=[100, 40], v=10.), duration=8.)
        self.trajs = [lead, w1, w2]
        self.X0s = [[0, 100, -np.pi, 0, 10], [-15, 0, np.pi / 2, 0, 10], [5, 0, np.pi / 2, 0, 10]]
        self.extends = (-30, 110, -10, 110)
        self.windfield = WindField([0, 2.5])
        self.time = np.arange(0., 17.5, _default_dt)
        super().__init__()


@register
class ScenPatrol3(Scenario):                                 # :230-248
    name, desc = "patrol_3", "dev patrol"

    def __init__(self, nv=2):
        self.trajs = []
        for i in range(nv):
            dy = 5 * i
            dx = dy / 2
            leg = ddt.TrajectoryLine([0, 10 + dy], [100 - dx, 10 + dy], v=10., t0=0.)
            turn = ddt.TrajectoryCircle(c=[100 - dx, 40], r=30. - dy, v=10., t0=0., alpha0=-np.pi / 2, dalpha=np.pi)
            back = ddtf.TrajSlalom(p1=[100, 60 - dy], p2=[0, 60 - dy], v=10., t0=0., phi=np.pi / 2)
            self.trajs.append(ddt.CompositeTraj([leg, turn, back]))
        self.windfield = WindField([0, 5.])
        super().__init__()


@register
class ScenCircularFormation(Scenario):                       # :254-264
    name, desc = "circForm", "circular formation"

    def __init__(self):
        P0s = [[30, 10], [40, 10]]
        self.trajs = [ddt.TrajectoryCircle(alpha0=3 * np.pi / 2 + i * np.pi / 6) for i in range(len(P0s))]
        super().__init__()
        for P0, X0 in zip(P0s, self.X0s):
            X0[:Aircraft.s_y + 1] = P0
        self.ppctl = False
        self.windfield = WindField([0, 5.])                  # set after X0s were derived with zero wind, as upstream


@register
class ScenOval(Scenario):                                    # :268-279
    """Upstream never calls Scenario.__init__ here, so `aircrafts` / `perts` are missing and test_simulation raises
    AttributeError; the defaults are filled in (one trajectory -> the first of the two X0s is used)."""
    name, desc = "oval", "oval"

    def __init__(self):
        self.trajs = [ddtf.TrajLineWithIntro(Y0=[0., 100.], Y1=[0., 50.], Y2=[200., 50.], r=25.)]
        self.X0s = [[0, 100, -np.pi, 0, 10], [0, 0, -np.pi, 0, 10]]
        self.extends = (-30, 110, -10, 110)
        self.windfield = WindField([0, 2.5])
        self.time = np.arange(0, 20, 0.01)
        super().__init__()


@register
class ScenDualOpty(Scenario):                                # :283-290 (planner outputs as tabulated references)
    name, desc = "dual opty", "dual opty"
    files = ("optyplan_exp6_0.npz", "optyplan_exp6_1.npz", "optyplan_exp6_2.npz")

    def __init__(self, files=None):
        self.trajs = [ddtf.TrajTabulated(f) for f in (files or self.files)]      # the .npz files are not shipped upstream either
        super().__init__()


@register
class ScenOpty2(ScenDualOpty):                               # :294-303
    name, desc = "opty2", "opty2"
    files = tuple(f"optyplan_exp7_{i}.npz" for i in range(5))


def print_available():
    print("Available scenarios:")
    for i, n in enumerate(list_available()):
        print(f"{i} -> {n}")


def get(_name):
    return _scenarios[_name][1](), _scenarios[_name][0]
