"""render() / query() on RAW parameters: the activations (softplus, bounded sigmoid / exp, normalize) run inside the
preprocess kernels and the per-Gaussian backward kernels return gradients with respect to the raw parameters
(libr2xray: r2x_*_raw).  The reference applies them as six small torch kernels per iteration and
differentiates through them with autograd (`r2_gaussian/gaussian/gaussian_model.py:112-126`).

Used by `render_query.render/query` when the model offers `raw_parameters()` (this repository's GaussianModel) and
`R2X_FUSED_ACTIVATIONS` is not 0; models that only expose `get_density / get_scaling / get_rotation` (e.g. the
reference's own class) keep the plain path.  Training mode only has the speculative forward (no host synchronisation;
`_C.speculative`), evaluation synchronises like the plain path.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _C
from ._lib import ActivationDesc, check, load


def enabled() -> bool:
    return os.environ.get("R2X_FUSED_ACTIVATIONS", "1") != "0"


def _act(params) -> ActivationDesc:
    a = ActivationDesc()
    bound = params.get("scale_bound")
    if bound is None:
        a.scale_mode, a.scale_lo, a.scale_hi = 0, 0.0, 0.0
    else:
        a.scale_mode, a.scale_lo, a.scale_hi = 1, float(bound[0]), float(bound[1])
    return a


def _forward_loop(run, key, P, per_gaussian, dev, training):
    """Provision the binning buffer, run the forward; training: speculative (capacity resolved in the backward),
    otherwise synchronise and grow on overflow.  Returns (NumRendered, binning)."""
    lib = load()
    cap = _C._Workspace.capacity(key, P, per_gaussian)
    spec = training and key in _C._Workspace.hints and os.environ.get("R2X_SPECULATIVE", "1") != "0"
    if spec:
        cap = _C._Workspace._round(max(2 * cap, per_gaussian * P))
    status = torch.empty(2, dtype=torch.int32, device=dev)
    while True:
        binning = torch.empty(lib.r2x_binning_bytes(cap), dtype=torch.uint8, device=dev)
        run(binning, cap, status)
        if spec:
            return _C._pending(status, cap, key, dev), binning
        R, overflow = status.tolist()
        _C._Workspace.update(key, R)
        if not overflow:
            return _C.NumRendered(R, cap), binning
        cap = _C._Workspace.capacity(key, P, per_gaussian)


class _RasterizeRaw(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, raw_density, raw_scales, raw_rotations, settings, act):
        lib = load()
        s = settings
        dev = means3D.device
        P, H, W = int(means3D.shape[0]), int(s.image_height), int(s.image_width)
        f = lambda t: _C._f32(t, dev)
        means3D, raw_density, raw_scales, raw_rotations = f(means3D), f(raw_density), f(raw_scales), f(raw_rotations)
        view, proj, campos = f(s.viewmatrix), f(s.projmatrix), f(s.campos)
        with torch.cuda.device(dev):
            u8 = dict(dtype=torch.uint8, device=dev)
            color = torch.empty((1, H, W), dtype=torch.float32, device=dev)
            radii = torch.empty((P,), dtype=torch.int32, device=dev)
            geom = torch.empty(lib.r2x_raster_geom_bytes(P), **u8)
            img = torch.empty(lib.r2x_raster_image_bytes(P, W, H), **u8)
            stream = torch.cuda.current_stream(dev).cuda_stream

            def run(binning, cap, status):
                rc = lib.r2x_raster_forward_async_raw(
                    stream, P, W, H, _C._ptr(means3D), _C._ptr(raw_density), _C._ptr(raw_scales), float(s.scale_modifier),
                    _C._ptr(raw_rotations), _C._ptr(view), _C._ptr(proj), _C._ptr(campos), float(s.tanfovx), float(s.tanfovy),
                    int(s.mode), color.data_ptr(), _C._ptr(radii), geom.data_ptr(), img.data_ptr(), binning.data_ptr(), cap,
                    status.data_ptr(), C.byref(act))
                check(rc, "r2x_raster_forward_async_raw")

            R, binning = _forward_loop(run, ("raster", dev.index, P, W, H), P, 12, dev, any(ctx.needs_input_grad))
        ctx.settings, ctx.act, ctx.num_rendered = s, act, R
        ctx.save_for_backward(means3D, raw_scales, raw_rotations, radii, geom, binning, img, view, proj, campos)
        ctx.mark_non_differentiable(radii)
        return color, radii

    @staticmethod
    def backward(ctx, grad_color, _grad_radii):
        lib = load()
        s, act, R = ctx.settings, ctx.act, ctx.num_rendered
        means3D, raw_scales, raw_rotations, radii, geom, binning, img, view, proj, campos = ctx.saved_tensors
        if getattr(R, "pending", None) is not None:
            R.resolve()
        cap = _C._carved_capacity(binning, R)
        dev = means3D.device
        P, H, W = int(means3D.shape[0]), int(s.image_height), int(s.image_width)
        with torch.cuda.device(dev):
            opts = dict(dtype=torch.float32, device=dev)
            g2 = torch.empty((P, 3), **opts); gd = torch.empty((P, 1), **opts); g3 = torch.empty((P, 3), **opts)
            gcov = torch.empty((P, 6), **opts); gs = torch.empty((P, 3), **opts); gr = torch.empty((P, 4), **opts)
            scratch = torch.empty(lib.r2x_raster_bwd_scratch_bytes(cap), dtype=torch.uint8, device=dev)
            dL = _C._f32(grad_color, dev)
            rc = lib.r2x_raster_backward_raw(
                torch.cuda.current_stream(dev).cuda_stream, P, cap, W, H, _C._ptr(means3D), _C._ptr(raw_scales),
                float(s.scale_modifier), _C._ptr(raw_rotations), _C._ptr(view), _C._ptr(proj), _C._ptr(campos),
                float(s.tanfovx), float(s.tanfovy), _C._ptr(radii), _C._ptr(geom), _C._ptr(binning), _C._ptr(img),
                scratch.data_ptr(), _C._ptr(dL), _C._ptr(g2), _C._ptr(gd), _C._ptr(g3), _C._ptr(gcov), _C._ptr(gs), _C._ptr(gr),
                int(s.mode), C.byref(act))
            check(rc, "r2x_raster_backward_raw")
        return g3, g2, gd, gs, gr, None, None


class _VoxelizeRaw(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, raw_density, raw_scales, raw_rotations, settings, act):
        lib = load()
        s = settings
        dev = means3D.device
        P = int(means3D.shape[0])
        nx, ny, nz = int(s.nVoxel_x), int(s.nVoxel_y), int(s.nVoxel_z)
        f = lambda t: _C._f32(t, dev)
        means3D, raw_density, raw_scales, raw_rotations = f(means3D), f(raw_density), f(raw_scales), f(raw_rotations)
        grid = (nx, ny, nz, float(s.sVoxel_x), float(s.sVoxel_y), float(s.sVoxel_z), float(s.center_x), float(s.center_y),
                float(s.center_z))
        with torch.cuda.device(dev):
            u8 = dict(dtype=torch.uint8, device=dev)
            vol = torch.empty((nx, ny, nz), dtype=torch.float32, device=dev)
            rx = torch.empty((P,), dtype=torch.int32, device=dev); ry = torch.empty_like(rx); rz = torch.empty_like(rx)
            geom = torch.empty(lib.r2x_voxel_geom_bytes(P), **u8)
            img = torch.empty(lib.r2x_voxel_image_bytes(P, nx, ny, nz), **u8)
            stream = torch.cuda.current_stream(dev).cuda_stream

            def run(binning, cap, status):
                rc = lib.r2x_voxel_forward_async_raw(
                    stream, P, *grid, _C._ptr(means3D), _C._ptr(raw_density), _C._ptr(raw_scales), float(s.scale_modifier),
                    _C._ptr(raw_rotations), vol.data_ptr(), _C._ptr(rx), _C._ptr(ry), _C._ptr(rz), geom.data_ptr(),
                    img.data_ptr(), binning.data_ptr(), cap, status.data_ptr(), C.byref(act))
                check(rc, "r2x_voxel_forward_async_raw")

            key = ("voxel", dev.index, P, nx, ny, nz, round(float(s.sVoxel_x) / nx, 6))
            R, binning = _forward_loop(run, key, P, 8, dev, any(ctx.needs_input_grad))
        ctx.settings, ctx.act, ctx.num_rendered, ctx.grid = s, act, R, grid
        ctx.save_for_backward(means3D, raw_scales, raw_rotations, rx, ry, rz, geom, binning, img)
        return vol, (rx, ry, rz)

    @staticmethod
    def backward(ctx, grad_vol, _grad_radii):
        lib = load()
        s, act, R, grid = ctx.settings, ctx.act, ctx.num_rendered, ctx.grid
        means3D, raw_scales, raw_rotations, rx, ry, rz, geom, binning, img = ctx.saved_tensors
        if getattr(R, "pending", None) is not None:
            R.resolve()
        cap = _C._carved_capacity(binning, R)
        dev = means3D.device
        P = int(means3D.shape[0])
        with torch.cuda.device(dev):
            opts = dict(dtype=torch.float32, device=dev)
            gd = torch.empty((P, 1), **opts); g3 = torch.empty((P, 3), **opts); gcov = torch.empty((P, 6), **opts)
            gs = torch.empty((P, 3), **opts); gr = torch.empty((P, 4), **opts)
            scratch = torch.empty(lib.r2x_voxel_bwd_scratch_bytes(cap), dtype=torch.uint8, device=dev)
            dL = _C._f32(grad_vol, dev)
            rc = lib.r2x_voxel_backward_raw(
                torch.cuda.current_stream(dev).cuda_stream, P, cap, *grid, _C._ptr(means3D), _C._ptr(raw_scales),
                float(s.scale_modifier), _C._ptr(raw_rotations), _C._ptr(rx), _C._ptr(ry), _C._ptr(rz), _C._ptr(geom),
                _C._ptr(binning), _C._ptr(img), scratch.data_ptr(), _C._ptr(dL), _C._ptr(gd), _C._ptr(g3), _C._ptr(gcov),
                _C._ptr(gs), _C._ptr(gr), C.byref(act))
            check(rc, "r2x_voxel_backward_raw")
        return g3, gd, gs, gr, None, None


def rasterize_raw(means3D, means2D, raw, settings):
    """raw = model.raw_parameters(): {"density", "scaling", "rotation", "scale_bound"} -> (image [1,H,W], radii)."""
    return _RasterizeRaw.apply(means3D, means2D, raw["density"], raw["scaling"], raw["rotation"], settings, _act(raw))


def voxelize_raw(means3D, raw, settings):
    return _VoxelizeRaw.apply(means3D, raw["density"], raw["scaling"], raw["rotation"], settings, _act(raw))
