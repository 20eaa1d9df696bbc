"""Stand-ins for third-party modules the reference's host code imports at module load but that are not installed in
this image (matplotlib, open3d, scikit-image, plyfile, SimpleITK).  They exist ONLY so that the reference's own
train.py / test.py can be run UNCHANGED on the drop-in packages (scripts/run_reference_drivers.py): importing works,
anything the drivers do not actually execute raises with a clear message.  Not part of the product package."""


class Absent:
    """Attribute access yields further placeholders; calling one raises."""

    def __init__(self, name):
        self.__dict__["_name"] = name

    def __getattr__(self, item):
        return Absent(f"{self._name}.{item}")

    def __call__(self, *a, **k):
        raise RuntimeError(f"{self._name} is not installed in this image (scripts/ref_shims placeholder)")

    def __mro_entries__(self, bases):   # allows `class X(placeholder)` at import time
        return (object,)
