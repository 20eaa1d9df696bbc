"""Persistent-workspace projector / voxelizer over the asynchronous C ABI.

`render()` / `query()` through the autograd wrappers follow the reference call for call (one host
round trip to size the binning buffer, fresh state buffers every call).  Throughput paths -- bench.py,
evaluation sweeps over many views, the Gaussian-sharded multi-GPU projector -- use these engines instead:
all state lives in buffers allocated once, the forward never synchronises with the host, and the
instance capacity is checked after the fact (`check()`), growing the workspace and re-running if a scene
ever needs more instances than provisioned.
"""
from __future__ import annotations

import torch

from ._lib import check, load


def _ptr(t):
    return None if t is None or t.numel() == 0 else t.data_ptr()


class RasterEngine:
    """Forward X-ray projector for a fixed (P, W, H) with preallocated state."""

    def __init__(self, P: int, W: int, H: int, device="cuda", capacity: int | None = None):
        self.lib = load()
        self.P, self.W, self.H = int(P), int(W), int(H)
        self.device = torch.device(device)
        with torch.cuda.device(self.device):
            u8 = dict(dtype=torch.uint8, device=self.device)
            self.geom = torch.empty(self.lib.r2x_raster_geom_bytes(self.P), **u8)
            self.img = torch.empty(self.lib.r2x_raster_image_bytes(self.P, self.W, self.H), **u8)
            self.radii = torch.empty(self.P, dtype=torch.int32, device=self.device)
            self.out = torch.empty((1, self.H, self.W), dtype=torch.float32, device=self.device)
            self.status = torch.zeros(2, dtype=torch.int32, device=self.device)
            self.capacity = 0
            self.binning = None
            self._reserve(capacity if capacity is not None else max(16 * self.P, 1 << 16))

    def _reserve(self, capacity: int):
        self.capacity = int(capacity)
        self.binning = torch.empty(self.lib.r2x_binning_bytes(self.capacity), dtype=torch.uint8, device=self.device)

    def forward(self, means, dens, scales, rots, viewmatrix, projmatrix, campos, tanfovx, tanfovy, mode,
                scale_modifier: float = 1.0, cov3D_precomp=None, out=None):
        """Enqueue one projection on the current stream; returns the [1,H,W] output tensor."""
        out = self.out if out is None else out
        rc = self.lib.r2x_raster_forward_async(
            torch.cuda.current_stream(self.device).cuda_stream, self.P, self.W, self.H, _ptr(means), _ptr(dens),
            _ptr(scales), float(scale_modifier), _ptr(rots), _ptr(cov3D_precomp), _ptr(viewmatrix), _ptr(projmatrix),
            _ptr(campos), float(tanfovx), float(tanfovy), 0, int(mode), out.data_ptr(), self.radii.data_ptr(),
            self.geom.data_ptr(), self.img.data_ptr(), self.binning.data_ptr(), self.capacity, self.status.data_ptr())
        check(rc, "r2x_raster_forward_async")
        return out

    def render_only(self, out=None):
        """Re-run only the per-tile accumulation kernel on the state of the last forward (profiling)."""
        out = self.out if out is None else out
        rc = self.lib.r2x_raster_render_only(torch.cuda.current_stream(self.device).cuda_stream, self.P, self.W, self.H,
                                             self.capacity, self.geom.data_ptr(), self.binning.data_ptr(),
                                             self.img.data_ptr(), out.data_ptr())
        check(rc, "r2x_raster_render_only")
        return out

    def num_rendered(self) -> int:
        """Synchronises; instance count of the last forward."""
        return int(self.status.cpu()[0].item())

    def check(self) -> bool:
        """Synchronises; True if the last forward fitted the capacity, else grows it (caller re-runs)."""
        R, ov = (int(v) for v in self.status.cpu().tolist())
        if ov:
            self._reserve(int(R * 1.25) + 1024)
            return False
        return True

    def fit(self, *fwd_args, **fwd_kw):
        """Run forward until it fits, then trim the capacity to 1.25x the need.  Returns R."""
        while True:
            self.forward(*fwd_args, **fwd_kw)
            if self.check():
                break
        R = self.num_rendered()
        return R


class VoxelEngine:
    """Forward density-volume query for a fixed (P, grid) with preallocated state."""

    def __init__(self, P: int, nVoxel, device="cuda", capacity: int | None = None):
        self.lib = load()
        self.P = int(P)
        self.nx, self.ny, self.nz = (int(v) for v in nVoxel)
        self.device = torch.device(device)
        with torch.cuda.device(self.device):
            u8 = dict(dtype=torch.uint8, device=self.device)
            self.geom = torch.empty(self.lib.r2x_voxel_geom_bytes(self.P), **u8)
            self.img = torch.empty(self.lib.r2x_voxel_image_bytes(self.P, self.nx, self.ny, self.nz), **u8)
            self.radii = torch.empty((3, self.P), dtype=torch.int32, device=self.device)
            self.out = torch.empty((self.nx, self.ny, self.nz), dtype=torch.float32, device=self.device)
            self.status = torch.zeros(2, dtype=torch.int32, device=self.device)
            self.capacity = 0
            self.binning = None
            self._reserve(capacity if capacity is not None else max(32 * self.P, 1 << 16))

    def _reserve(self, capacity: int):
        self.capacity = int(capacity)
        self.binning = torch.empty(self.lib.r2x_binning_bytes(self.capacity), dtype=torch.uint8, device=self.device)

    def forward(self, means, dens, scales, rots, sVoxel, center, scale_modifier: float = 1.0, cov3D_precomp=None,
                out=None):
        out = self.out if out is None else out
        rc = self.lib.r2x_voxel_forward_async(
            torch.cuda.current_stream(self.device).cuda_stream, self.P, self.nx, self.ny, self.nz, float(sVoxel[0]),
            float(sVoxel[1]), float(sVoxel[2]), float(center[0]), float(center[1]), float(center[2]), _ptr(means),
            _ptr(dens), _ptr(scales), float(scale_modifier), _ptr(rots), _ptr(cov3D_precomp), 0, out.data_ptr(),
            self.radii[0].data_ptr(), self.radii[1].data_ptr(), self.radii[2].data_ptr(), self.geom.data_ptr(),
            self.img.data_ptr(), self.binning.data_ptr(), self.capacity, self.status.data_ptr())
        check(rc, "r2x_voxel_forward_async")
        return out

    def render_only(self, out=None):
        out = self.out if out is None else out
        rc = self.lib.r2x_voxel_render_only(torch.cuda.current_stream(self.device).cuda_stream, self.P, self.nx, self.ny,
                                            self.nz, self.capacity, self.geom.data_ptr(), self.binning.data_ptr(),
                                            self.img.data_ptr(), out.data_ptr())
        check(rc, "r2x_voxel_render_only")
        return out

    def num_rendered(self) -> int:
        return int(self.status.cpu()[0].item())

    def check(self) -> bool:
        R, ov = (int(v) for v in self.status.cpu().tolist())
        if ov:
            self._reserve(int(R * 1.25) + 1024)
            return False
        return True

    def fit(self, *fwd_args, **fwd_kw):
        while True:
            self.forward(*fwd_args, **fwd_kw)
            if self.check():
                break
        return self.num_rendered()
