"""GPU parity AT THE BASELINE CONFIGURATIONS (BASELINE.json configs / SURVEY.md 8d): our kernels, through the
reference-shaped `_C` entry points, against the UNMODIFIED reference CUDA sources compiled into
oracle/_ref/libr2ref.so -- forward and backward.

  raster  100k Gaussians / 512^2 cone beam   (headline; init-like and trained-like cloud, 3 views each)
  raster   50k Gaussians / 256^2             (config 1)
  raster  300k Gaussians / 512^2             (config 3, the cloud the 8 shards are cut from)
  voxel   256^3 over 500k Gaussians          (config 4; RAS/forward.cu:198-395, VOX/forward.cu:58-315)
  voxel   32^3 TV crop over 100k Gaussians   (train.py:128-139)
  + the binding paths cov3D_precomp (PYX/rasterization.py:241-253), markVisible
    (SUB/rasterize_points.cu:166-186) and debug=True (synchronous ABI, allocator callback).

Bars: radii / tiles_touched / multiset of 64-bit (tile | depth bits) keys / per-tile ranges bit-exact;
intensities |ours - ref| <= 1e-5 * max|ref| + 1e-7; gradients |a-b| <= 5e-4 |b| + 5e-5 max|b| (the reference's
float atomics are not order-stable, util.grad_mismatch)."""
import numpy as np
import pytest

import util
from r2_gaussian_b200 import scene

pytestmark = pytest.mark.gpu

GRAD_KEYS_R = ["dL_dmean2D", "dL_dopacity", "dL_dmu", "dL_dmean3D", "dL_dcov3D", "dL_dscale", "dL_drot"]
GRAD_KEYS_V = ["dL_dopacity", "dL_dmean3D", "dL_dcov3D", "dL_dscale", "dL_drot"]
_clouds = {}


def cloud_of(P, kind):
    key = (P, kind)
    if key not in _clouds:
        _clouds.clear()     # keep one big cloud alive at a time
        _clouds[key] = scene.make_cloud(P, kind=kind, seed=0)
    return _clouds[key]


def image_bar(ours, ref, what):
    scale = float(np.abs(ref).max())
    err = float(np.abs(ours.astype(np.float64) - ref.astype(np.float64)).max())
    assert err <= 1e-5 * scale + 1e-7, f"{what}: max |ours - ref| = {err:.3g} vs scale {scale:.3g}"
    return err / max(scale, 1e-30)


def check_raster(cloud, view, backward=True, cov3D_precomp=None, debug=False):
    dL = np.random.RandomState(5).randn(view.image_height, view.image_width).astype(np.float32) if backward else None
    ref = util.run_ref_raster(cloud, view, dL, cov3D_precomp=cov3D_precomp)
    ours = util.ours_raster_forward(cloud, view, cov3D_precomp=cov3D_precomp, debug=debug)
    assert int(ours["R"]) == ref["R"]
    np.testing.assert_array_equal(ours["radii"], ref["radii"])
    np.testing.assert_array_equal(ours["tiles_touched"], ref["tiles_touched"])
    vis = ref["radii"] > 0
    np.testing.assert_array_equal(ours["depth"][vis].view(np.uint32), ref["depth"][vis].view(np.uint32))
    np.testing.assert_array_equal(ours["xy"][vis].view(np.uint32), ref["xy"][vis].view(np.uint32))
    np.testing.assert_array_equal(ours["conic_opacity"][vis].view(np.uint32), ref["conic_opacity"][vis].view(np.uint32))
    np.testing.assert_array_equal(ours["mu"][vis].view(np.uint32), ref["mu"][vis].view(np.uint32))
    assert util.key_multiset_equal(ours["keys"], ref["keys"])
    np.testing.assert_array_equal(ours["ranges"], ref["ranges"])
    image_bar(ours["image"], ref["image"], "image")
    if backward:
        g = util.ours_raster_backward(cloud, view, ours, dL, debug=debug)
        keys = GRAD_KEYS_R if cov3D_precomp is None else ["dL_dmean2D", "dL_dopacity", "dL_dmu", "dL_dmean3D", "dL_dcov3D"]
        if cov3D_precomp is None:
            util.assert_grads_close(g, ref["grads"], keys, rtol=5e-4, atol_rel=5e-5, label="ours vs ref ")
        else:
            for k in keys:
                assert util.grad_mismatch(g[k], ref["grads"][k], 5e-4, 5e-5) <= 1.0, k
            assert not g["dL_dscale"].any() and not g["dL_drot"].any()
    return ours, ref


@pytest.mark.parametrize("kind", ["init", "trained"])
def test_raster_100k_512_cone_three_views(kind):
    """The headline scene of BASELINE.json (100k Gaussians, 512^2 cone beam)."""
    cloud = cloud_of(100_000, kind)
    sc = scene.cone_beam_scanner(512, 256)
    for angle in (0.0, 2.1, 4.4):
        check_raster(cloud, scene.make_view(sc, angle))


@pytest.mark.parametrize("kind", ["init", "trained"])
def test_raster_50k_256_cone(kind):
    cloud = cloud_of(50_000, kind)
    sc = scene.cone_beam_scanner(256, 256)
    for angle in (0.3, 3.9):
        check_raster(cloud, scene.make_view(sc, angle))


def test_raster_300k_512_cone():
    cloud = cloud_of(300_000, "trained")
    sc = scene.cone_beam_scanner(512, 256)
    check_raster(cloud, scene.make_view(sc, 1.0))


def test_raster_100k_512_parallel():
    cloud = cloud_of(100_000, "trained")
    sc = scene.parallel_beam_scanner(512, 256)
    check_raster(cloud, scene.make_view(sc, 0.7))


def check_voxel(cloud, nV, sV, ctr, backward=True, cov3D_precomp=None, debug=False):
    dL = np.random.RandomState(6).randn(*nV).astype(np.float32) if backward else None
    ref = util.run_ref_voxel(cloud, nV, sV, ctr, dL, cov3D_precomp=cov3D_precomp)
    ours = util.ours_voxel_forward(cloud, nV, sV, ctr, cov3D_precomp=cov3D_precomp, debug=debug)
    assert int(ours["R"]) == ref["R"]
    for k in ["radii_x", "radii_y", "radii_z", "tiles_touched"]:
        np.testing.assert_array_equal(ours[k], ref[k])
    vis = ref["tiles_touched"] > 0
    np.testing.assert_array_equal(ours["xyz_vol"][vis].view(np.uint32), ref["xyz_vol"][vis].view(np.uint32))
    np.testing.assert_array_equal(ours["depth"][vis].view(np.uint32), ref["depth"][vis].view(np.uint32))
    assert util.key_multiset_equal(ours["keys"], ref["keys"])
    np.testing.assert_array_equal(ours["ranges"], ref["ranges"])
    image_bar(ours["vol"], ref["vol"], "volume")
    if backward:
        g = util.ours_voxel_backward(cloud, nV, sV, ctr, ours, dL, debug=debug)
        if cov3D_precomp is None:
            util.assert_grads_close(g, ref["grads"], GRAD_KEYS_V, rtol=5e-4, atol_rel=5e-5, label="ours vs ref ")
        else:
            for k in ["dL_dopacity", "dL_dmean3D", "dL_dcov3D"]:
                assert util.grad_mismatch(g[k], ref["grads"][k], 5e-4, 5e-5) <= 1.0, k
    return ours, ref


@pytest.mark.parametrize("kind", ["init", "trained"])
def test_voxel_256_cube_500k(kind):
    """BASELINE config 4: 256^3 volume query over 500k Gaussians (32768 tiles)."""
    cloud = cloud_of(500_000, kind)
    check_voxel(cloud, (256, 256, 256), (2.0, 2.0, 2.0), (0.0, 0.0, 0.0), backward=(kind == "trained"))


def test_voxel_tv_crop_32_100k():
    """The 32^3 total-variation crop of train.py:128-139 (random centre, crop = volume / 8 per axis)."""
    cloud = cloud_of(100_000, "trained")
    rng = np.random.RandomState(3)
    sV = (0.25, 0.25, 0.25)
    for _ in range(3):
        ctr = tuple(float(v) for v in (2.0 - 0.25) * (rng.rand(3) - 0.5))
        check_voxel(cloud, (32, 32, 32), sV, ctr)


# ---- binding paths -------------------------------------------------------------------------------
def _precomputed_cov(cloud, view):
    """Sigma3 per Gaussian, bit-identical to what the reference's preprocess computes from scales / rotations
    (the oracle is pinned to it bit for bit, test_ref_gpu.py)."""
    return util.oracle_raster_forward(cloud, view, render=False)["cov3D"]


def test_raster_cov3D_precomp_path():
    cloud = cloud_of(20_000, "trained")
    view = scene.make_view(scene.cone_beam_scanner(256, 256), 0.9)
    cov = _precomputed_cov(cloud, view)
    ours_c, ref_c = check_raster(cloud, view, cov3D_precomp=cov)
    ours_s, _ = check_raster(cloud, view, backward=False)
    np.testing.assert_array_equal(ours_c["radii"], ours_s["radii"])
    np.testing.assert_array_equal(ours_c["image"].view(np.uint32), ours_s["image"].view(np.uint32))


def test_voxel_cov3D_precomp_path():
    """Forward against the reference.  The reference's voxelizer backward cannot run with cov3D_precomp (it needs
    `scales` for the radius, VOX/forward.cu:137, and then differentiates through `rotations`, which are absent:
    VOX/backward.cu:201-211 dereferences a null pointer), so the gradients are checked against the CPU oracle."""
    cloud = cloud_of(20_000, "trained")
    view = scene.make_view(scene.cone_beam_scanner(64, 64), 0.0)
    cov = _precomputed_cov(cloud, view)
    grid = ((64, 64, 64), (2.0, 2.0, 2.0), (0.0, 0.0, 0.0))
    ours, _ = check_voxel(cloud, *grid, backward=False, cov3D_precomp=cov)
    ours_s, _ = check_voxel(cloud, *grid, backward=False)
    np.testing.assert_array_equal(ours["vol"].view(np.uint32), ours_s["vol"].view(np.uint32))
    dL = np.random.RandomState(6).randn(*grid[0]).astype(np.float32)
    g = util.ours_voxel_backward(cloud, *grid, ours, dL)
    orc = util.oracle_voxel_forward(cloud, *grid, cov3D_precomp=cov)
    from oracle import r2_oracle as o
    go = o.voxel_backward(orc, cloud.scales, None, grid[0], grid[1], dL, cov3D_precomp=cov)
    util.assert_grads_close(g, go, ["dL_dopacity", "dL_dmean3D", "dL_dcov3D"])
    assert not g["dL_dscale"].any() and not g["dL_drot"].any()


def test_debug_true_synchronous_abi():
    """debug=True takes r2x_raster_forward / r2x_voxel_forward: the variant with the reference's semantics exactly
    (the library learns R, asks the allocator callback for the binning buffer, syncs after every stage)."""
    cloud = cloud_of(20_000, "trained")
    view = scene.make_view(scene.cone_beam_scanner(256, 256), 2.2)
    a, _ = check_raster(cloud, view, debug=True)
    b, _ = check_raster(cloud, view, backward=False, debug=False)
    np.testing.assert_array_equal(a["image"].view(np.uint32), b["image"].view(np.uint32))
    assert int(a["R"]) == a["R"].capacity          # the synchronous ABI sizes the buffer exactly
    check_voxel(cloud, (64, 64, 64), (2.0, 2.0, 2.0), (0.0, 0.0, 0.0), debug=True)


def test_mark_visible_matches_reference():
    import ctypes as C

    import torch
    from r2_gaussian_b200.rasterization import GaussianRasterizationSettings, GaussianRasterizer

    lib = util.need_ref()
    cloud = cloud_of(20_000, "trained")
    view = scene.make_view(scene.cone_beam_scanner(256, 256), 1.3)
    t = util.to_torch(cloud, view)
    # move a third of the points behind / onto the near plane (view-space z <= 0.2 fails, RAS/auxiliary.h:143-168)
    means = t["means"].clone()
    cam = torch.tensor(view.campos, device="cuda")
    means[::3] = cam + (means[::3] - cam) * torch.linspace(-0.02, 0.06, means[::3].shape[0], device="cuda")[:, None]
    settings = GaussianRasterizationSettings(view.image_height, view.image_width, view.tanfovx, view.tanfovy, 1.0,
                                             t["view"], t["proj"], t["campos"], False, view.mode, False)
    ours = GaussianRasterizer(settings).markVisible(means)
    present = torch.zeros(cloud.P, dtype=torch.bool, device="cuda")
    p = lambda x: C.c_void_p(x.data_ptr())
    lib.ref_mark_visible(cloud.P, p(means), p(t["view"]), p(t["proj"]), p(present))
    torch.cuda.synchronize()
    assert ours.dtype == torch.bool and ours.shape == (cloud.P,)
    assert 0 < int(present.sum()) < cloud.P
    assert torch.equal(ours, present)
    from oracle import r2_oracle as orc
    np.testing.assert_array_equal(orc.mark_visible(means.cpu().numpy(), view.viewmatrix, view.projmatrix),
                                  present.cpu().numpy())
