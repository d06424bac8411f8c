"""Import-only stand-in for matplotlib so the reference's d2d modules load
headless (golden-vector generation only; never used by product code)."""
import sys, types

class _Anything:
    def __getattr__(self, name): return _Anything()
    def __call__(self, *a, **k): return _Anything()
    def __iter__(self): return iter((_Anything(), _Anything()))   # `fig, ax = plt.subplots()`
    def __getitem__(self, k): return _Anything()
    def __len__(self): return 2

def _mk(name):
    m = types.ModuleType(name)
    m.__getattr__ = lambda attr: _Anything()
    sys.modules[name] = m
    return m

for _sub in ("pyplot", "animation", "image", "offsetbox", "transforms", "patches", "cm"):
    setattr(sys.modules[__name__], _sub, _mk(f"matplotlib.{_sub}"))
