from _absent import Absent as _A


def use(*a, **k):   # matplotlib.use("Agg") at import time is harmless
    return None


def __getattr__(name):
    return _A(f"matplotlib.{name}")
