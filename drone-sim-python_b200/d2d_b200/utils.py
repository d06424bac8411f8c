"""Import-compatible stand-in for `d2d.utils` (d2d/utils.py:7-18): the angle wrap is evaluated by the engine, the wind
field is the constant one of the planners with this module's (x, y, t) argument order."""
from .guidance import norm_mpi_pi  # noqa: F401  (d2d/utils.py:7)
from . import opty_utils as _d2ou


class WindField(_d2ou.WindField):
    """d2d/utils.py:10-18 samples with (x, y, t); d2d/opty_utils.py:18-27 with (t, x, y).  The wind is constant, so both
    return the stored vector; only the signatures differ."""

    def _constant(self, *_position_and_time):
        return self.w

    sample_num = sample_sym = _constant
