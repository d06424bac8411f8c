from . import direct_collocation
