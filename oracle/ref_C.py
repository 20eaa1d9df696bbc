"""`_C` stand-in that drives the UNMODIFIED reference CUDA kernels (oracle/_ref/libr2ref.so) with the argument lists of
the reference's pybind module (SUB/ext.cpp:17-23, SUB/rasterize_points.h:18-61, SUB/voxelize_points.cu:29-167).

TEST / BASELINE INFRASTRUCTURE ONLY -- never imported by the product package.  scripts/run_reference_drivers.py places
a copy of the reference's own Python package (PYX/{__init__,rasterization,voxelization}.py, unchanged) next to a
one-line `_C.py` that re-exports this module, so the reference's train.py / test.py can be run on the reference's own
kernels without building its 10-minute torch extension; the result is the yardstick for the same drivers running on
the drop-in packages.  The three scratch buffers live inside the shim library (oracle/ref_shim.cu) and belong to the
most recent forward, which is exactly how the drivers use them (one render + one query per iteration, each followed
by its own backward); the byte tensors returned here are placeholders."""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def _load():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "_ref", "libr2ref.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: run `bash oracle/build_ref.sh` where /root/reference exists")
        _lib = C.CDLL(path)
        _lib.ref_raster_forward.restype = C.c_int
        _lib.ref_voxel_forward.restype = C.c_int
    return _lib


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None and t.numel() else None


def _c(t, dev):
    if t.numel() == 0:
        return t
    return t.to(device=dev, dtype=torch.float32).contiguous()


_f = C.c_float


def _join_default_stream():
    """The reference kernels run on the legacy default stream (torch's default stream is that stream)."""
    st = torch.cuda.current_stream()
    if st.cuda_stream != 0:
        st.synchronize()


def rasterize_gaussians(means3D, opacity, scales, rotations, scale_modifier, cov3D_precomp, viewmatrix, projmatrix,
                        tan_fovx, tan_fovy, image_height, image_width, campos, prefiltered, mode, debug):
    lib = _load()
    dev = means3D.device
    P, H, W = int(means3D.shape[0]), int(image_height), int(image_width)
    out = torch.zeros((1, H, W), dtype=torch.float32, device=dev)
    radii = torch.zeros((P,), dtype=torch.int32, device=dev)
    R = 0
    if P:
        a = [_c(t, dev) for t in (means3D, opacity, scales, rotations, cov3D_precomp, viewmatrix, projmatrix, campos)]
        with torch.cuda.device(dev):
            _join_default_stream()
            R = lib.ref_raster_forward(P, W, H, _p(a[0]), _p(a[1]), _p(a[2]), _f(scale_modifier), _p(a[3]), _p(a[4]),
                                       _p(a[5]), _p(a[6]), _p(a[7]), _f(tan_fovx), _f(tan_fovy), int(mode), _p(out),
                                       _p(radii))
    e = torch.empty(0, dtype=torch.uint8, device=dev)
    return R, out, radii, e, e.clone(), e.clone()


def rasterize_gaussians_backward(means3D, radii, scales, rotations, scale_modifier, cov3D_precomp, viewmatrix,
                                 projmatrix, tan_fovx, tan_fovy, dL_dout_color, campos, geomBuffer, R, binningBuffer,
                                 imageBuffer, mode, debug):
    lib = _load()
    dev = means3D.device
    P = int(means3D.shape[0])
    H, W = int(dL_dout_color.shape[-2]), int(dL_dout_color.shape[-1])
    z = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)
    g2, gc, go, gm, g3, gcov, gs, gr = z(P, 3), z(P, 2, 2), z(P, 1), z(P, 1), z(P, 3), z(P, 6), z(P, 3), z(P, 4)
    if P:
        a = [_c(t, dev) for t in (means3D, scales, rotations, cov3D_precomp, viewmatrix, projmatrix, campos,
                                  dL_dout_color)]
        with torch.cuda.device(dev):
            _join_default_stream()
            lib.ref_raster_backward(P, int(R), W, H, _p(a[0]), _p(a[1]), _f(scale_modifier), _p(a[2]), _p(a[3]),
                                    _p(a[4]), _p(a[5]), _p(a[6]), _f(tan_fovx), _f(tan_fovy), _p(radii), _p(a[7]),
                                    _p(g2), _p(gc), _p(go), _p(gm), _p(g3), _p(gcov), _p(gs), _p(gr), int(mode))
    return g2, go, gm, g3, gcov, gs, gr


def mark_visible(means3D, viewmatrix, projmatrix):
    lib = _load()
    dev = means3D.device
    P = int(means3D.shape[0])
    present = torch.zeros((P,), dtype=torch.bool, device=dev)
    if P:
        a = [_c(t, dev) for t in (means3D, viewmatrix, projmatrix)]
        with torch.cuda.device(dev):
            _join_default_stream()
            lib.ref_mark_visible(P, _p(a[0]), _p(a[1]), _p(a[2]), _p(present))
    return present


def voxelize_gaussians(means3D, opacity, scales, rotations, scale_modifier, cov3D_precomp, nVoxel_x, nVoxel_y,
                       nVoxel_z, sVoxel_x, sVoxel_y, sVoxel_z, center_x, center_y, center_z, prefiltered, debug):
    lib = _load()
    dev = means3D.device
    P = int(means3D.shape[0])
    nx, ny, nz = int(nVoxel_x), int(nVoxel_y), int(nVoxel_z)
    vol = torch.zeros((nx, ny, nz), dtype=torch.float32, device=dev)
    rx = torch.zeros((P,), dtype=torch.int32, device=dev)
    ry, rz = torch.zeros_like(rx), torch.zeros_like(rx)
    R = 0
    if P:
        a = [_c(t, dev) for t in (means3D, opacity, scales, rotations, cov3D_precomp)]
        with torch.cuda.device(dev):
            _join_default_stream()
            R = lib.ref_voxel_forward(P, nx, ny, nz, _f(sVoxel_x), _f(sVoxel_y), _f(sVoxel_z), _f(center_x),
                                      _f(center_y), _f(center_z), _p(a[0]), _p(a[1]), _p(a[2]), _f(scale_modifier),
                                      _p(a[3]), _p(a[4]), _p(vol), _p(rx), _p(ry), _p(rz))
    e = torch.empty(0, dtype=torch.uint8, device=dev)
    return R, vol, rx, ry, rz, e, e.clone(), e.clone()


def voxelize_gaussians_backward(means3D, radii_x, radii_y, radii_z, scales, rotations, scale_modifier, cov3D_precomp,
                                dL_dout, geomBuffer, R, binningBuffer, imageBuffer, nVoxel_x, nVoxel_y, nVoxel_z,
                                sVoxel_x, sVoxel_y, sVoxel_z, center_x, center_y, center_z, debug):
    lib = _load()
    dev = means3D.device
    P = int(means3D.shape[0])
    z = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)
    gn, gc6, go, g3, gcov, gs, gr = z(P, 3), z(P, 6), z(P, 1), z(P, 3), z(P, 6), z(P, 3), z(P, 4)
    if P:
        a = [_c(t, dev) for t in (means3D, scales, rotations, cov3D_precomp, dL_dout)]
        with torch.cuda.device(dev):
            _join_default_stream()
            lib.ref_voxel_backward(P, int(R), int(nVoxel_x), int(nVoxel_y), int(nVoxel_z), _f(sVoxel_x), _f(sVoxel_y),
                                   _f(sVoxel_z), _f(center_x), _f(center_y), _f(center_z), _p(a[0]), _p(a[1]),
                                   _f(scale_modifier), _p(a[2]), _p(a[3]), _p(radii_x), _p(radii_y), _p(radii_z),
                                   _p(a[4]), _p(gn), _p(gc6), _p(go), _p(g3), _p(gcov), _p(gs), _p(gr))
    return go, g3, gcov, gs, gr
