from _absent import Absent as _A


def __getattr__(name):
    return _A(f"open3d.{name}")
