"""One training iteration as a fixed sequence of libr2xray calls on preallocated buffers.

The reference's iteration (train.py:104-160) is render() -> L1 + D-SSIM -> query() of a random TV crop -> backward ->
densification statistics -> Adam, expressed as ~270 torch kernels behind autograd.  With the fused kernels of this
package the GPU needs ~0.55 ms for it at 100k Gaussians / 512^2, but the autograd path still spends ~1 ms of HOST time
per iteration (about fifty small torch ops, four autograd.Function round trips, thirty allocations).  `NativeTrainStep`
issues the same kernels directly:

    raster forward (raw parameters)      r2x_raster_forward_async_raw        [image summed over ranks when sharded]
    L1 + D-SSIM value and gradient       r2x_image_loss
    TV-crop query forward                r2x_voxel_forward_async_raw         [volume summed over ranks when sharded]
    TV value and gradient                r2x_tv3d_loss
    both backward passes                 r2x_voxel_backward_raw, r2x_raster_backward_raw
    densification statistics             r2x_densify_stats
    Adam on the four parameter tensors   r2x_adam_step_sum  (gradient = raster part + voxel part)

No autograd graph, no per-iteration allocation, no host synchronisation: both forwards are speculative (instance
capacity provisioned from the previous call of the same shape, `_C._Workspace`), and the statistics / Adam launches
are GUARDED by the forwards' overflow flags on the device, so an overflowed iteration changes nothing; the host reads
the flags one iteration late (`check()`), raises the capacity and repeats that iteration.

The model's tensors are updated in place: `GaussianModel._xyz/_density/_scaling/_rotation`, the `FusedAdam` state of
`gaussians.optimizer` (same `exp_avg`, `exp_avg_sq`, `step`, so checkpoints and the densification surgery are
unchanged) and `max_radii2D / xyz_gradient_accum / denom`.  After densification (new tensors) the step re-binds itself.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import torch

from . import _C, fused, sharded
from ._lib import AdamGroup, check, load


def enabled() -> bool:
    return fused.enabled() and os.environ.get("R2X_NATIVE_STEP", "1") != "0"


class NativeTrainStep:
    def __init__(self, gaussians, lambda_dssim: float, lambda_tv: float = 0.0, tv_vol_nVoxel=None, tv_vol_sVoxel=None,
                 scaling_modifier: float = 1.0):
        self.gm = gaussians
        self.lambda_dssim, self.lambda_tv = float(lambda_dssim), float(lambda_tv)
        self.tv_n = None if tv_vol_nVoxel is None else tuple(int(v) for v in tv_vol_nVoxel)
        self.tv_s = None if tv_vol_sVoxel is None else tuple(float(v) for v in tv_vol_sVoxel)
        self.use_tv = self.lambda_tv > 0 and self.tv_n is not None
        self.scale_modifier = float(scaling_modifier)
        self.lib = load()
        self._bound = None          # identity of the tensors the buffers were made for
        self._pending = None        # (host status words, event, keys, caps, args) of the last enqueued iteration
        self.repeats = 0            # iterations repeated after a capacity overflow

    # ------------------------------------------------------------------ buffers
    def _signature(self, H, W):
        gm = self.gm
        return (sharded.enabled(), gm._xyz.data_ptr(), gm._density.data_ptr(), gm._scaling.data_ptr(), gm._rotation.data_ptr(),
                int(gm._xyz.shape[0]), H, W, gm.max_radii2D.data_ptr(), gm.xyz_gradient_accum.data_ptr())

    def _bind(self, H, W):
        gm, lib = self.gm, self.lib
        dev = gm._xyz.device
        P = int(gm._xyz.shape[0])
        self.P, self.H, self.W, self.dev = P, H, W, dev
        f32 = dict(dtype=torch.float32, device=dev)
        u8 = dict(dtype=torch.uint8, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            # raster.  Gaussian-sharded runs: the image travels through the exchange with one extra word, this rank's
            # overflow flag, so that after the sum EVERY rank knows whether ANY rank's forward overflowed (the summed
            # image is then wrong everywhere) and all ranks skip / repeat the iteration together.
            self.sharded = sharded.enabled()
            self.image_ext = torch.zeros(H * W + 4, **f32)
            self.image = self.image_ext[:H * W].view(1, H, W)
            self.flag_r = self.image_ext[H * W:H * W + 1]
            self.radii = torch.empty((P,), **i32)
            self.geom = torch.empty(lib.r2x_raster_geom_bytes(P), **u8)
            self.img = torch.empty(lib.r2x_raster_image_bytes(P, W, H), **u8)
            self.status_r = torch.zeros(2, **i32)
            self.key_r = ("raster", dev.index, P, W, H)
            self.g2 = torch.empty((P, 3), **f32); self.gd = torch.empty((P, 1), **f32); self.g3 = torch.empty((P, 3), **f32)
            self.gcov = torch.empty((P, 6), **f32); self.gs = torch.empty((P, 3), **f32); self.gr = torch.empty((P, 4), **f32)
            # image loss
            self.loss_scratch_bytes = int(lib.r2x_image_loss_scratch_bytes(H, W))
            self.loss_scratch = torch.empty(self.loss_scratch_bytes, **u8)
            self.loss_out = torch.zeros(3, **f32)            # L1, SSIM, lambda_l1 L1 + lambda_dssim (1 - SSIM)
            self.dL_dimage = torch.empty((H, W), **f32)
            # TV crop
            if self.use_tv:
                nx, ny, nz = self.tv_n
                self.vol_ext = torch.zeros(nx * ny * nz + 4, **f32)
                self.vol = self.vol_ext[:nx * ny * nz].view(nx, ny, nz)
                self.flag_v = self.vol_ext[nx * ny * nz:nx * ny * nz + 1]
                self.rx = torch.empty((P,), **i32); self.ry = torch.empty((P,), **i32); self.rz = torch.empty((P,), **i32)
                self.geom_v = torch.empty(lib.r2x_voxel_geom_bytes(P), **u8)
                self.img_v = torch.empty(lib.r2x_voxel_image_bytes(P, nx, ny, nz), **u8)
                self.status_v = torch.zeros(2, **i32)
                self.key_v = ("voxel", dev.index, P, nx, ny, nz, round(self.tv_s[0] / nx, 6))
                self.tv_scratch_bytes = int(lib.r2x_tv3d_scratch_bytes(nx, ny, nz))
                self.tv_scratch = torch.empty(self.tv_scratch_bytes, **u8)
                self.tv_out = torch.zeros(1, **f32)
                self.dL_dvol = torch.empty((nx, ny, nz), **f32)
                self.gdv = torch.empty((P, 1), **f32); self.g3v = torch.empty((P, 3), **f32); self.gcovv = torch.empty((P, 6), **f32)
                self.gsv = torch.empty((P, 3), **f32); self.grv = torch.empty((P, 4), **f32)
            self.cap_r = self.cap_v = 0
            self.binning_r = self.binning_v = self.scratch_r = self.scratch_v = None
        # Adam: the optimizer's own state tensors, one group per parameter tensor (xyz, density, scaling, rotation)
        opt = gm.optimizer
        self.adam = []
        grads = {id(gm._xyz): (self.g3, self.g3v if self.use_tv else None),
                 id(gm._density): (self.gd, self.gdv if self.use_tv else None),
                 id(gm._scaling): (self.gs, self.gsv if self.use_tv else None),
                 id(gm._rotation): (self.gr, self.grv if self.use_tv else None)}
        for group in opt.param_groups:
            for p in group["params"]:
                if id(p) not in grads:
                    raise RuntimeError("NativeTrainStep: the optimizer holds a parameter that is not one of the model's four")
                st = opt.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                self.adam.append((group, p, st, *grads[id(p)]))
        if len(self.adam) != 4:
            raise RuntimeError("NativeTrainStep: expected the four parameter groups of GaussianModel.training_setup")
        n = len(self.adam)
        self.adam_groups = (AdamGroup * n)()
        self.adam_grads2 = (C.c_void_p * n)()
        for k, (group, p, st, g1, g2) in enumerate(self.adam):
            a = self.adam_groups[k]
            a.param, a.grad, a.exp_avg, a.exp_avg_sq = p.data_ptr(), g1.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
            a.numel = p.numel()
            self.adam_grads2[k] = g2.data_ptr() if g2 is not None else None
        self.act = fused._act(gm.raw_parameters())
        self._bound = self._signature(H, W)

    def _provision(self):
        """Instance capacities for this iteration (generous: an overflow needs the count to double between two calls)."""
        lib, dev = self.lib, self.dev
        W_ = _C._Workspace
        want_r = W_._round(max(2 * W_.capacity(self.key_r, self.P, 12), 12 * self.P))
        if want_r > self.cap_r:
            self.cap_r = want_r
            self.binning_r = torch.empty(lib.r2x_binning_bytes(want_r), dtype=torch.uint8, device=dev)
            self.scratch_r = torch.empty(lib.r2x_raster_bwd_scratch_bytes(want_r), dtype=torch.uint8, device=dev)
        if self.use_tv:
            want_v = W_._round(max(2 * W_.capacity(self.key_v, self.P, 8), 8 * self.P))
            if want_v > self.cap_v:
                self.cap_v = want_v
                self.binning_v = torch.empty(lib.r2x_binning_bytes(want_v), dtype=torch.uint8, device=dev)
                self.scratch_v = torch.empty(lib.r2x_voxel_bwd_scratch_bytes(want_v), dtype=torch.uint8, device=dev)

    # ------------------------------------------------------------------ one iteration
    def __call__(self, cam, gt, tv_centre=None, apply_update: bool = True):
        """Enqueue one iteration.  `cam`: camera (render_query.render's contract); `gt`: [1,H,W] or [H,W] CUDA float32
        target; `tv_centre`: 3 floats (ignored without TV).  Returns {"render", "radii", "loss" (device [3]: L1, SSIM,
        image total), "tv" (device [1] or None)} -- views of buffers that the next call overwrites."""
        self.check()                                            # the previous iteration (one late; no stall)
        H, W = int(cam.image_height), int(cam.image_width)
        if self._bound != self._signature(H, W):
            self._bind(H, W)
        if self.P == 0 and not self.sharded:
            raise RuntimeError("NativeTrainStep: empty model")
        # an EMPTY SHARD of a Gaussian-sharded run goes through the same sequence: the library calls are no-ops that
        # produce a zero image / volume (and zero status words), and the rank takes part in both exchanges with the same
        # buffer sizes as its peers
        args = (cam, gt, None if tv_centre is None else tuple(float(v) for v in tv_centre), bool(apply_update))
        self._enqueue(*args)
        return self.result

    def _enqueue(self, cam, gt, tv_centre, apply_update):
        gm, lib, dev, P, H, W = self.gm, self.lib, self.dev, self.P, self.H, self.W
        self._provision()
        mode = int(cam.mode)
        if mode == 0:
            tfx = tfy = 1.0
        elif mode == 1:
            tfx, tfy = math.tan(cam.FoVx * 0.5), math.tan(cam.FoVy * 0.5)
        else:
            raise ValueError("Unsupported mode!")
        gt = gt.reshape(H, W)
        if gt.dtype != torch.float32 or not gt.is_contiguous() or gt.device != dev:
            gt = gt.to(device=dev, dtype=torch.float32).contiguous()
        view, proj, campos = cam.world_view_transform, cam.full_proj_transform, cam.camera_center
        act = C.byref(self.act)
        sm = self.scale_modifier
        xyz, dens, scal, rot = gm._xyz, gm._density, gm._scaling, gm._rotation
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream(dev).cuda_stream
            check(lib.r2x_raster_forward_async_raw(
                st, P, W, H, xyz.data_ptr(), dens.data_ptr(), scal.data_ptr(), sm, rot.data_ptr(), view.data_ptr(),
                proj.data_ptr(), campos.data_ptr(), tfx, tfy, mode, self.image.data_ptr(), self.radii.data_ptr(),
                self.geom.data_ptr(), self.img.data_ptr(), self.binning_r.data_ptr(), self.cap_r, self.status_r.data_ptr(),
                act), "r2x_raster_forward_async_raw")
            image = self.image
            if self.sharded:
                self.flag_r.copy_(self.status_r[1:2])          # my overflow flag rides with the image
                sharded.sharded_sum_(self.image_ext)
                self.status_r[1:2].copy_(self.flag_r)          # ... and comes back as "any rank overflowed"
            check(lib.r2x_image_loss(st, H, W, image.data_ptr(), gt.data_ptr(), 1.0, self.lambda_dssim,
                                     self.loss_out.data_ptr(), self.dL_dimage.data_ptr(), self.loss_scratch.data_ptr(),
                                     self.loss_scratch_bytes), "r2x_image_loss")
            if self.use_tv:
                nx, ny, nz = self.tv_n
                grid = (nx, ny, nz, self.tv_s[0], self.tv_s[1], self.tv_s[2], tv_centre[0], tv_centre[1], tv_centre[2])
                check(lib.r2x_voxel_forward_async_raw(
                    st, P, *grid, xyz.data_ptr(), dens.data_ptr(), scal.data_ptr(), sm, rot.data_ptr(), self.vol.data_ptr(),
                    self.rx.data_ptr(), self.ry.data_ptr(), self.rz.data_ptr(), self.geom_v.data_ptr(), self.img_v.data_ptr(),
                    self.binning_v.data_ptr(), self.cap_v, self.status_v.data_ptr(), act), "r2x_voxel_forward_async_raw")
                vol = self.vol
                if self.sharded:
                    self.flag_v.copy_(self.status_v[1:2])
                    sharded.sharded_sum_(self.vol_ext)
                    self.status_v[1:2].copy_(self.flag_v)
                check(lib.r2x_tv3d_loss(st, nx, ny, nz, vol.data_ptr(), 1, self.tv_out.data_ptr(), self.dL_dvol.data_ptr(),
                                        self.tv_scratch.data_ptr(), self.tv_scratch_bytes), "r2x_tv3d_loss")
                self.dL_dvol.mul_(self.lambda_tv)
                check(lib.r2x_voxel_backward_raw(
                    st, P, self.cap_v, *grid, xyz.data_ptr(), scal.data_ptr(), sm, rot.data_ptr(), self.rx.data_ptr(),
                    self.ry.data_ptr(), self.rz.data_ptr(), self.geom_v.data_ptr(), self.binning_v.data_ptr(),
                    self.img_v.data_ptr(), self.scratch_v.data_ptr(), self.dL_dvol.data_ptr(), self.gdv.data_ptr(),
                    self.g3v.data_ptr(), self.gcovv.data_ptr(), self.gsv.data_ptr(), self.grv.data_ptr(), act),
                    "r2x_voxel_backward_raw")
            check(lib.r2x_raster_backward_raw(
                st, P, self.cap_r, W, H, xyz.data_ptr(), scal.data_ptr(), sm, rot.data_ptr(), view.data_ptr(), proj.data_ptr(),
                campos.data_ptr(), tfx, tfy, self.radii.data_ptr(), self.geom.data_ptr(), self.binning_r.data_ptr(),
                self.img.data_ptr(), self.scratch_r.data_ptr(), self.dL_dimage.data_ptr(), self.g2.data_ptr(),
                self.gd.data_ptr(), self.g3.data_ptr(), self.gcov.data_ptr(), self.gs.data_ptr(), self.gr.data_ptr(), mode,
                act), "r2x_raster_backward_raw")
            guard_v = self.status_v.data_ptr() if self.use_tv else None
            check(lib.r2x_densify_stats(st, P, self.radii.data_ptr(), self.g2.data_ptr(), gm.max_radii2D.data_ptr(),
                                        gm.xyz_gradient_accum.data_ptr(), gm.denom.data_ptr(), self.status_r.data_ptr(),
                                        guard_v), "r2x_densify_stats")
            if apply_update:
                steps = set()
                for k, (group, p, stt, _g1, _g2) in enumerate(self.adam):
                    stt["step"] += 1
                    steps.add(int(stt["step"].item()))
                    self.adam_groups[k].lr = float(group["lr"])
                if len(steps) != 1:
                    raise RuntimeError("NativeTrainStep: the four parameters' Adam step counts differ")
                b1, b2 = self.adam[0][0]["betas"]
                check(lib.r2x_adam_step_sum(st, len(self.adam), C.cast(self.adam_groups, C.c_void_p),
                                            C.cast(self.adam_grads2, C.c_void_p) if self.use_tv else None, float(b1),
                                            float(b2), float(self.adam[0][0]["eps"]), steps.pop(), self.status_r.data_ptr(),
                                            guard_v), "r2x_adam_step_sum")
            # the status words travel to pinned host memory behind an event; read one iteration late
            host = _C._Workspace.pinned_status(), (_C._Workspace.pinned_status() if self.use_tv else None)
            host[0].copy_(self.status_r, non_blocking=True)
            if self.use_tv:
                host[1].copy_(self.status_v, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(dev))
        self._pending = (host, ev, (cam, gt, tv_centre, apply_update))
        self.result = {"render": image, "radii": self.radii, "viewspace_grad": self.g2, "loss": self.loss_out,
                       "tv": self.tv_out if self.use_tv else None}

    def total_loss(self) -> float:
        """Host value of the last iteration's loss (synchronises; logging only)."""
        t = float(self.loss_out[2])
        return t + self.lambda_tv * float(self.tv_out[0]) if self.use_tv else t

    def check(self):
        """Resolve the last enqueued iteration: update the capacity hints; if a forward had overflowed (its launches
        changed nothing on the device), undo the host-side step count and run that iteration again."""
        if self._pending is None:
            return
        host, ev, args = self._pending
        self._pending = None
        ev.synchronize()
        Rr, ov_r = int(host[0][0]), int(host[0][1])
        _C._Workspace.update(self.key_r, Rr)
        _C._Workspace.release(host[0])
        ov_v = 0
        if self.use_tv:
            Rv, ov_v = int(host[1][0]), int(host[1][1])
            _C._Workspace.update(self.key_v, Rv)
            _C._Workspace.release(host[1])
        if ov_r or ov_v:
            self.repeats += 1
            if self.repeats > 8:
                raise _C.CapacityOverflow("NativeTrainStep: the instance capacity keeps overflowing")
            if args[3]:
                for _group, _p, stt, _g1, _g2 in self.adam:
                    stt["step"] -= 1
            self._enqueue(*args)
            self.check()

    flush = check
