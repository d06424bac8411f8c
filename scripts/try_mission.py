import sys, time as _t, numpy as np
sys.path.insert(0, "drone-sim-python_b200")
from d2d_b200 import mission
C4 = np.array([[0, -20], [25, -20], [25, -100], [0, -100]], float)
X1_f = ((0, 40, 0, 0, 12), (25, 40, 0, 0, 12), (25, -40, 0, 0, 12), (0, -40, 0, 0, 12))
t0 = _t.time()
X, U, _, _, Ur, eth, t, t_f = mission.CircularFormationGVF(C4, 60, 15, 4, X1_f, 0, 0.05, 1000)
print("phase 1 rows", len(X), "t_f", t_f, "wall", _t.time() - t0)
print("last states", X[-1])
print("e_theta last", eth[-1], "Ur last", Ur[-1])
# distance to criterion over time
d = np.abs(X[:, :, :3] - np.asarray(X1_f)[None, :, :3])
ok = (d <= np.array([3, 3, np.deg2rad(0.5)])).all(axis=2)
print("steps where each aircraft meets its own criterion:", ok.sum(0), " all together:", ok.all(1).sum())
