class Problem:
    def __init__(self, *a, **k):
        raise RuntimeError("opty is not installed; stub only lets d2d.opty_utils import")
