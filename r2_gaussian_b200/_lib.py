"""ctypes binding of libr2xray.so (C ABI in include/r2x.h).

There is NO fallback: if the shared library is missing or does not export every symbol the header
declares, importing the compute path raises.  `R2X_AUTOBUILD=1` (default) compiles it with nvcc when
the in-tree library is absent or stale.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libr2xray.so")

ALLOC_FN = C.CFUNCTYPE(C.c_void_p, C.c_size_t, C.c_void_p)

_vp, _i, _f, _ll, _sz = C.c_void_p, C.c_int, C.c_float, C.c_longlong, C.c_size_t

# name -> (restype, argtypes); kept in the order of include/r2x.h
PROTOTYPES = {
    "r2x_last_error": (C.c_char_p, []),
    "r2x_version": (_i, []),
    "r2x_raster_geom_bytes": (_sz, [_i]),
    "r2x_raster_image_bytes": (_sz, [_i, _i, _i]),
    "r2x_voxel_geom_bytes": (_sz, [_i]),
    "r2x_voxel_image_bytes": (_sz, [_i, _i, _i, _i]),
    "r2x_binning_bytes": (_sz, [_ll]),
    "r2x_raster_bwd_scratch_bytes": (_sz, [_ll]),
    "r2x_voxel_bwd_scratch_bytes": (_sz, [_ll]),
    "r2x_raster_forward": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _f, _f, _i, _i,
                                _vp, _vp, _vp, _vp, ALLOC_FN, _vp, _i, C.POINTER(_i)]),
    "r2x_raster_forward_async": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _f, _f, _i, _i,
                                      _vp, _vp, _vp, _vp, _vp, _ll, _vp]),
    "r2x_raster_backward": (_i, [_vp, _i, _ll, _i, _i, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _f, _f, _vp,
                                 _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i]),
    "r2x_raster_render_only": (_i, [_vp, _i, _i, _i, _ll, _vp, _vp, _vp, _vp]),
    "r2x_voxel_render_only": (_i, [_vp, _i, _i, _i, _i, _ll, _vp, _vp, _vp, _vp]),
    "r2x_mark_visible": (_i, [_vp, _i, _vp, _vp, _vp, _vp]),
    "r2x_raster_export": (_i, [_vp, _i, _i, _i, _ll, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "r2x_voxel_forward": (_i, [_vp, _i, _i, _i, _i, _f, _f, _f, _f, _f, _f, _vp, _vp, _vp, _f, _vp, _vp, _i,
                               _vp, _vp, _vp, _vp, _vp, _vp, ALLOC_FN, _vp, _i, C.POINTER(_i)]),
    "r2x_voxel_forward_async": (_i, [_vp, _i, _i, _i, _i, _f, _f, _f, _f, _f, _f, _vp, _vp, _vp, _f, _vp, _vp, _i,
                                     _vp, _vp, _vp, _vp, _vp, _vp, _vp, _ll, _vp]),
    "r2x_voxel_backward": (_i, [_vp, _i, _ll, _i, _i, _i, _f, _f, _f, _f, _f, _f, _vp, _vp, _f, _vp, _vp,
                                _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i]),
    "r2x_voxel_export": (_i, [_vp, _i, _i, _i, _i, _ll, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "r2x_knn_scratch_bytes": (_sz, [_i]),
    "r2x_knn3_mean_dist2": (_i, [_vp, _i, _vp, _vp, _vp, _sz]),
    "r2x_image_loss_scratch_bytes": (_sz, [_i, _i]),
    "r2x_image_loss": (_i, [_vp, _i, _i, _vp, _vp, _f, _f, _vp, _vp, _vp, _sz]),
    "r2x_tv3d_scratch_bytes": (_sz, [_i, _i, _i]),
    "r2x_tv3d_loss": (_i, [_vp, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _sz]),
    "r2x_adam_step": (_i, [_vp, _i, _vp, C.c_double, C.c_double, C.c_double, _ll]),
    "r2x_adam_step_sum": (_i, [_vp, _i, _vp, _vp, C.c_double, C.c_double, C.c_double, _ll, _vp, _vp]),
    "r2x_densify_stats": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "r2x_raster_forward_async_raw": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _f, _vp, _vp, _vp, _vp, _f, _f, _i, _vp, _vp, _vp, _vp,
                                          _vp, _ll, _vp, _vp]),
    "r2x_raster_backward_raw": (_i, [_vp, _i, _ll, _i, _i, _vp, _vp, _f, _vp, _vp, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp, _vp,
                                     _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "r2x_voxel_forward_async_raw": (_i, [_vp, _i, _i, _i, _i, _f, _f, _f, _f, _f, _f, _vp, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp,
                                         _vp, _vp, _vp, _ll, _vp, _vp]),
    "r2x_voxel_backward_raw": (_i, [_vp, _i, _ll, _i, _i, _i, _f, _f, _f, _f, _f, _f, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _vp,
                                    _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "r2x_mask_select_scratch_bytes": (_sz, [_i]),
    "r2x_mask_select": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _sz]),
    "r2x_gather_rows": (_i, [_vp, _i, _vp, _vp, _ll]),
    "r2x_peer_alloc": (_i, [_sz, C.POINTER(_vp)]),
    "r2x_peer_free": (_i, [_vp]),
    "r2x_ipc_export": (_i, [_vp, _vp]),
    "r2x_ipc_open": (_i, [_vp, C.POINTER(_vp)]),
    "r2x_ipc_close": (_i, [_vp]),
    "r2x_peer_allreduce_sum": (_i, [_vp, _i, _i, _vp, _vp, C.c_uint32, _vp, _ll, _vp]),
    "r2x_peer_allreduce_sum_t": (_i, [_vp, _i, _i, _vp, _vp, C.c_uint32, _vp, _ll, _vp, _ll]),
}


class ActivationDesc(C.Structure):
    """Mirror of `r2x_activation` (include/r2x.h)."""
    _fields_ = [("scale_mode", C.c_int), ("scale_lo", C.c_float), ("scale_hi", C.c_float)]


class GatherDesc(C.Structure):
    """Mirror of `r2x_gather_desc` (include/r2x.h)."""
    _fields_ = [("src0", C.c_void_p), ("src1", C.c_void_p), ("dst", C.c_void_p), ("n0", C.c_longlong), ("width", C.c_int)]


class AdamGroup(C.Structure):
    """Mirror of `r2x_adam_group` (include/r2x.h)."""
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("numel", C.c_longlong), ("lr", C.c_float)]

_lock = threading.Lock()
_lib = None


class R2XError(RuntimeError):
    pass


def load(autobuild: bool | None = None):
    """Load (building first if allowed and needed) and return the ctypes library handle."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if autobuild is None:
            autobuild = os.environ.get("R2X_AUTOBUILD", "1") != "0"
        if autobuild:
            from . import build as _build
            try:
                if _build.needs_build():
                    _build.build()
            except Exception as e:  # no nvcc on this box: use the prebuilt library if there is one
                if not os.path.exists(LIB_PATH):
                    raise R2XError(f"libr2xray.so is missing and could not be built: {e}") from e
        if not os.path.exists(LIB_PATH):
            raise R2XError(
                f"{LIB_PATH} not found. Build it with `python -m r2_gaussian_b200.build` "
                "(needs nvcc); there is no CPU fallback."
            )
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            try:
                fn = getattr(lib, name)
            except AttributeError as e:
                raise R2XError(f"libr2xray.so does not export {name}; rebuild it") from e
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().r2x_last_error()
        raise R2XError(f"{what} failed (code {rc}): {msg.decode() if msg else '?'}")
