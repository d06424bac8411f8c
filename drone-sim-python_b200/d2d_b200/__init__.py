"""d2d_b200 -- B200-native engine behind the `d2d` call surface of drone-sim-python.

Layout mirrors the reference package (`dynamic`, `guidance`, `trajectory`, `trajectory_factory`, `scenario`,
`opty_utils`, `multiopty_utils`) plus `simulation` (the loops of 05_test_simulation.py / 08_CircularFormation_Full.py),
`planner` (the Planner classes of 06_optyplan.py / 07_multioptyplan.py) and `engine` (the ctypes view of
include/d2dx.h).  Importing the package loads libd2dx.so; it raises when the library has not been built."""
from . import _lib  # noqa: F401  (fails loudly when libd2dx.so is missing)
from .engine import Engine, get_engine  # noqa: F401

__all__ = ["Engine", "get_engine"]
