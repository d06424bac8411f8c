"""Small helpers of `d2d.utils` (d2d/utils.py:7-18) kept for import compatibility."""
from .guidance import norm_mpi_pi  # noqa: F401  (d2d/utils.py:7; evaluated by the engine)


class WindField:
    """Constant wind with the numeric / symbolic sampling names of d2d/utils.py:10-18."""

    def __init__(self, w=[0., 0.]):
        self.w = w

    def sample_num(self, _x, _y, _t):
        return self.w

    def sample_sym(self, _x, _y, _t):
        return self.w
