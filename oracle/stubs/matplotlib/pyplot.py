"""Stub for `matplotlib.pyplot`."""


def __getattr__(name):
    def _f(*a, **k):
        raise RuntimeError("matplotlib stub")
    return _f
