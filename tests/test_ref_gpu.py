"""Pins the CPU oracle (and our kernels) against the UNMODIFIED reference CUDA sources compiled into
oracle/_ref/libr2ref.so (oracle/build_ref.sh) and run on the GPU: bit-exact radii / tiles_touched /
sorted 64-bit keys / point lists, 1e-5 on intensities, gradient agreement within atomics noise."""
import numpy as np
import pytest

import util
from r2_gaussian_b200 import scene

pytestmark = pytest.mark.gpu


run_ref_raster = util.run_ref_raster
run_ref_voxel = util.run_ref_voxel


@pytest.mark.parametrize("name", ["cone_trained_small", "parallel_trained_small", "cone_trained_ragged", "cone_init_mid"])
def test_oracle_pinned_to_reference_raster(name):
    cloud, view = util.case(name)
    dL = np.random.RandomState(5).randn(view.image_height, view.image_width).astype(np.float32)
    ref = run_ref_raster(cloud, view, dL)
    orc = util.oracle_raster_forward(cloud, view)
    assert ref["R"] == orc["R"]
    np.testing.assert_array_equal(ref["radii"], orc["radii"])
    np.testing.assert_array_equal(ref["tiles_touched"], orc["tiles_touched"])
    vis = orc["radii"] > 0
    np.testing.assert_array_equal(ref["depth"][vis].view(np.uint32), orc["depth"][vis].view(np.uint32))
    np.testing.assert_array_equal(ref["xy"][vis].view(np.uint32), orc["xy"][vis].view(np.uint32))
    np.testing.assert_array_equal(ref["cov3D"][vis].view(np.uint32), orc["cov3D"][vis].view(np.uint32))
    np.testing.assert_array_equal(ref["conic_opacity"][vis].view(np.uint32), orc["conic_opacity"][vis].view(np.uint32))
    np.testing.assert_array_equal(ref["mu"][vis].view(np.uint32), orc["mu"][vis].view(np.uint32))
    np.testing.assert_array_equal(ref["keys"], orc["keys"])          # sorted order too
    np.testing.assert_array_equal(ref["point_list"], orc["point_list"])
    np.testing.assert_array_equal(ref["ranges"], orc["ranges"])
    scale = float(np.abs(ref["image"]).max())
    assert np.abs(ref["image"].astype(np.float64) - orc["image"]).max() <= 1e-5 * scale + 1e-7
    go = util.oracle_raster_backward(cloud, view, orc, dL)
    util.assert_grads_close(go, ref["grads"], list(ref["grads"]), rtol=5e-4, atol_rel=5e-5, label="oracle vs ref ")


@pytest.mark.parametrize("name", ["cone_trained_small", "cone_trained_mid"])
def test_ours_against_reference_raster(name):
    cloud, view = util.case(name)
    dL = np.random.RandomState(5).randn(view.image_height, view.image_width).astype(np.float32)
    ref = run_ref_raster(cloud, view, dL)
    ours = util.ours_raster_forward(cloud, view)
    assert ours["R"] == ref["R"]
    np.testing.assert_array_equal(ours["radii"], ref["radii"])
    np.testing.assert_array_equal(ours["tiles_touched"], ref["tiles_touched"])
    assert util.key_multiset_equal(ours["keys"], ref["keys"])
    np.testing.assert_array_equal(ours["ranges"], ref["ranges"])
    scale = float(np.abs(ref["image"]).max())
    assert np.abs(ours["image"].astype(np.float64) - ref["image"]).max() <= 1e-5 * scale + 1e-7
    g = util.ours_raster_backward(cloud, view, ours, dL)
    util.assert_grads_close(g, ref["grads"], list(ref["grads"]), rtol=5e-4, atol_rel=5e-5, label="ours vs ref ")


@pytest.mark.parametrize("grid", [((32, 32, 32), (2.0, 2.0, 2.0), (0.0, 0.0, 0.0)),
                                  ((20, 36, 28), (1.3, 2.0, 1.7), (0.1, -0.05, 0.2))])
def test_oracle_and_ours_against_reference_voxel(grid):
    nV, sV, ctr = grid
    cloud = scene.make_cloud(1500, kind="trained", seed=9)
    dL = np.random.RandomState(6).randn(*nV).astype(np.float32)
    ref = run_ref_voxel(cloud, nV, sV, ctr, dL)
    orc = util.oracle_voxel_forward(cloud, nV, sV, ctr)
    ours = util.ours_voxel_forward(cloud, nV, sV, ctr)
    for other in (orc, ours):
        assert other["R"] == ref["R"]
        for k in ["radii_x", "radii_y", "radii_z", "tiles_touched"]:
            np.testing.assert_array_equal(other[k], ref[k])
        assert util.key_multiset_equal(other["keys"], ref["keys"])
        np.testing.assert_array_equal(other["ranges"], ref["ranges"])
        scale = float(np.abs(ref["vol"]).max())
        assert np.abs(other["vol"].astype(np.float64) - ref["vol"]).max() <= 1e-5 * scale + 1e-7
    np.testing.assert_array_equal(orc["keys"], ref["keys"])
    np.testing.assert_array_equal(orc["point_list"], ref["point_list"])
    vis = ref["tiles_touched"] > 0
    np.testing.assert_array_equal(orc["conic_opacity"][vis].view(np.uint32), ref["conic_opacity"][vis].view(np.uint32))
    go = util.oracle_voxel_backward(cloud, nV, sV, orc, dL)
    g = util.ours_voxel_backward(cloud, nV, sV, ctr, ours, dL)
    util.assert_grads_close(go, ref["grads"], list(ref["grads"]), rtol=5e-4, atol_rel=5e-5, label="oracle vs ref ")
    util.assert_grads_close(g, ref["grads"], list(ref["grads"]), rtol=5e-4, atol_rel=5e-5, label="ours vs ref ")
