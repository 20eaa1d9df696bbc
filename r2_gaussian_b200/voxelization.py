"""Density-volume (voxelizer) front-end with the reference's Python surface.

Mirrors PYX/voxelization.py of the reference: `GaussianVoxelizationSettings` (:26-38),
`GaussianVoxelizer` (:228-266) and the autograd bridge `_VoxelizeGaussians` (:59-225): same names,
argument order, return values `(vol[nx,ny,nz], (radii_x, radii_y, radii_z))` and gradient order.
"""
from __future__ import annotations

from typing import NamedTuple

import torch
from torch import nn

from . import _C
from ._snapshot import call_with_snapshot
from .rasterization import _exactly_one_covariance_source


class GaussianVoxelizationSettings(NamedTuple):
    scale_modifier: float
    nVoxel_x: int
    nVoxel_y: int
    nVoxel_z: int
    sVoxel_x: float
    sVoxel_y: float
    sVoxel_z: float
    center_x: float
    center_y: float
    center_z: float
    prefiltered: bool
    debug: bool


class _VoxelizeGaussians(torch.autograd.Function):
    """forward inputs:  (means3D, opacities, scales, rotations, cov3Ds_precomp, settings)
    backward outputs: (d means3D, d opacities, d scales, d rotations, d cov3Ds_precomp, None)."""

    @staticmethod
    def forward(ctx, means3D, opacities, scales, rotations, cov3Ds_precomp, voxel_settings):
        s = voxel_settings
        native_args = (means3D, opacities, scales, rotations, s.scale_modifier, cov3Ds_precomp, s.nVoxel_x,
                       s.nVoxel_y, s.nVoxel_z, s.sVoxel_x, s.sVoxel_y, s.sVoxel_z, s.center_x, s.center_y,
                       s.center_z, s.prefiltered, s.debug)
        with _C.speculative(any(ctx.needs_input_grad) and not s.debug):
            num_rendered, vol, rx, ry, rz, geom, binning, img = call_with_snapshot(
                _C.voxelize_gaussians, native_args, s.debug, "snapshot_fw.dump", "forward")
        ctx.voxel_settings = s
        ctx.num_rendered = num_rendered
        ctx.save_for_backward(means3D, scales, rotations, cov3Ds_precomp, rx, ry, rz, geom, binning, img)
        return vol, (rx, ry, rz)

    @staticmethod
    def backward(ctx, grad_vol, _grad_radii):
        s = ctx.voxel_settings
        means3D, scales, rotations, cov3Ds_precomp, rx, ry, rz, geom, binning, img = ctx.saved_tensors
        native_args = (means3D, rx, ry, rz, scales, rotations, s.scale_modifier, cov3Ds_precomp, grad_vol, geom,
                       ctx.num_rendered, binning, img, s.nVoxel_x, s.nVoxel_y, s.nVoxel_z, s.sVoxel_x, s.sVoxel_y,
                       s.sVoxel_z, s.center_x, s.center_y, s.center_z, s.debug)
        g_opac, g_means3D, g_cov, g_scales, g_rots = call_with_snapshot(
            _C.voxelize_gaussians_backward, native_args, s.debug, "snapshot_bw.dump", "backward")
        return g_means3D, g_opac, g_scales, g_rots, g_cov, None


def voxelize_gaussians(means3D, opacities, scales, rotations, cov3Ds_precomp, voxel_settings):
    return _VoxelizeGaussians.apply(means3D, opacities, scales, rotations, cov3Ds_precomp, voxel_settings)


class GaussianVoxelizer(nn.Module):
    def __init__(self, voxel_settings: GaussianVoxelizationSettings):
        super().__init__()
        self.voxel_settings = voxel_settings

    def forward(self, means3D, opacities, scales=None, rotations=None, cov3D_precomp=None):
        _exactly_one_covariance_source(scales, rotations, cov3D_precomp)
        empty = torch.Tensor([])
        return voxelize_gaussians(
            means3D, opacities,
            empty if scales is None else scales,
            empty if rotations is None else rotations,
            empty if cov3D_precomp is None else cov3D_precomp,
            self.voxel_settings)
