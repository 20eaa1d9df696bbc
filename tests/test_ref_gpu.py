"""Pins the CPU oracle (and our kernels) against the UNMODIFIED reference CUDA sources compiled into
oracle/_ref/libr2ref.so (oracle/build_ref.sh) and run on the GPU: bit-exact radii / tiles_touched /
sorted 64-bit keys / point lists, 1e-5 on intensities, gradient agreement within atomics noise."""
import ctypes as C

import numpy as np
import pytest

import util
from r2_gaussian_b200 import scene

pytestmark = pytest.mark.gpu


def _need_ref():
    lib = util.ref_lib()
    if lib is None:
        pytest.skip("oracle/_ref/libr2ref.so not built (needs /root/reference at build time)")
    return lib


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None and t.numel() else None


def run_ref_raster(cloud, view, dL=None):
    import torch

    lib = _need_ref()
    t = util.to_torch(cloud, view)
    P, W, H = cloud.P, view.image_width, view.image_height
    dev = "cuda"
    out = torch.zeros((1, H, W), device=dev); radii = torch.zeros(P, dtype=torch.int32, device=dev)
    f = C.c_float
    R = lib.ref_raster_forward(P, W, H, _ptr(t["means"]), _ptr(t["dens"]), _ptr(t["scales"]), f(1.0), _ptr(t["rots"]),
                               None, _ptr(t["view"]), _ptr(t["proj"]), _ptr(t["campos"]), f(view.tanfovx),
                               f(view.tanfovy), view.mode, _ptr(out), _ptr(radii))
    torch.cuda.synchronize()
    T = ((W + 15) // 16) * ((H + 15) // 16)
    depth = torch.zeros(P, device=dev); xy = torch.zeros((P, 2), device=dev); cov = torch.zeros((P, 6), device=dev)
    co = torch.zeros((P, 4), device=dev); mu = torch.zeros(P, device=dev)
    tt = torch.zeros(P, dtype=torch.int32, device=dev); po = torch.zeros(P, dtype=torch.int32, device=dev)
    ks = torch.zeros(max(R, 1), dtype=torch.int64, device=dev); pl = torch.zeros(max(R, 1), dtype=torch.int32, device=dev)
    rg = torch.zeros((T, 2), dtype=torch.int32, device=dev); nc = torch.zeros((H, W), dtype=torch.int32, device=dev)
    lib.ref_raster_export(P, W, H, R, _ptr(depth), _ptr(xy), _ptr(cov), _ptr(co), _ptr(mu), _ptr(tt), _ptr(po), None,
                          None, _ptr(ks), _ptr(pl), _ptr(rg), _ptr(nc))
    res = dict(R=R, image=out[0].cpu().numpy(), radii=radii.cpu().numpy(), depth=depth.cpu().numpy(),
               xy=xy.cpu().numpy(), cov3D=cov.cpu().numpy(), conic_opacity=co.cpu().numpy(), mu=mu.cpu().numpy(),
               tiles_touched=tt.cpu().numpy().astype(np.uint32), keys=ks.cpu().numpy().astype(np.uint64)[:R],
               point_list=pl.cpu().numpy().astype(np.uint32)[:R], ranges=rg.cpu().numpy().astype(np.uint32))
    if dL is not None:
        z = lambda *s: torch.zeros(s, device=dev)
        g2, gc, go, gm, g3, gcov, gs, gr = z(P, 3), z(P, 4), z(P, 1), z(P, 1), z(P, 3), z(P, 6), z(P, 3), z(P, 4)
        dLt = torch.tensor(dL, device=dev)
        lib.ref_raster_backward(P, R, W, H, _ptr(t["means"]), _ptr(t["scales"]), f(1.0), _ptr(t["rots"]), None,
                                _ptr(t["view"]), _ptr(t["proj"]), _ptr(t["campos"]), f(view.tanfovx), f(view.tanfovy),
                                _ptr(radii), _ptr(dLt), _ptr(g2), _ptr(gc), _ptr(go), _ptr(gm), _ptr(g3), _ptr(gcov),
                                _ptr(gs), _ptr(gr), view.mode)
        torch.cuda.synchronize()
        res["grads"] = dict(dL_dmean2D=g2.cpu().numpy(), dL_dopacity=go.cpu().numpy(), dL_dmu=gm.cpu().numpy(),
                            dL_dmean3D=g3.cpu().numpy(), dL_dcov3D=gcov.cpu().numpy(), dL_dscale=gs.cpu().numpy(),
                            dL_drot=gr.cpu().numpy())
    return res


def run_ref_voxel(cloud, nV, sV, ctr, dL=None):
    import torch

    lib = _need_ref()
    t = util.to_torch(cloud, None)
    P = cloud.P
    nx, ny, nz = nV
    dev = "cuda"
    f = C.c_float
    vol = torch.zeros(nV, device=dev)
    rx = torch.zeros(P, dtype=torch.int32, device=dev); ry = torch.zeros_like(rx); rz = torch.zeros_like(rx)
    R = lib.ref_voxel_forward(P, nx, ny, nz, f(sV[0]), f(sV[1]), f(sV[2]), f(ctr[0]), f(ctr[1]), f(ctr[2]),
                              _ptr(t["means"]), _ptr(t["dens"]), _ptr(t["scales"]), f(1.0), _ptr(t["rots"]), None,
                              _ptr(vol), _ptr(rx), _ptr(ry), _ptr(rz))
    torch.cuda.synchronize()
    T = ((nx + 7) // 8) * ((ny + 7) // 8) * ((nz + 7) // 8)
    depth = torch.zeros(P, device=dev); xyz = torch.zeros((P, 3), device=dev); cov = torch.zeros((P, 6), device=dev)
    co = torch.zeros((P, 7), device=dev)
    tt = torch.zeros(P, dtype=torch.int32, device=dev); po = torch.zeros(P, dtype=torch.int32, device=dev)
    ks = torch.zeros(max(R, 1), dtype=torch.int64, device=dev); pl = torch.zeros(max(R, 1), dtype=torch.int32, device=dev)
    rg = torch.zeros((T, 2), dtype=torch.int32, device=dev)
    lib.ref_voxel_export(P, nx, ny, nz, R, _ptr(depth), _ptr(xyz), _ptr(cov), _ptr(co), _ptr(tt), _ptr(po), None, None,
                         _ptr(ks), _ptr(pl), _ptr(rg), None)
    res = dict(R=R, vol=vol.cpu().numpy(), radii_x=rx.cpu().numpy(), radii_y=ry.cpu().numpy(), radii_z=rz.cpu().numpy(),
               depth=depth.cpu().numpy(), xyz_vol=xyz.cpu().numpy(), conic_opacity=co.cpu().numpy(),
               tiles_touched=tt.cpu().numpy().astype(np.uint32), keys=ks.cpu().numpy().astype(np.uint64)[:R],
               point_list=pl.cpu().numpy().astype(np.uint32)[:R], ranges=rg.cpu().numpy().astype(np.uint32))
    if dL is not None:
        z = lambda *s: torch.zeros(s, device=dev)
        gn, gc, go, g3, gcov, gs, gr = z(P, 3), z(P, 6), z(P, 1), z(P, 3), z(P, 6), z(P, 3), z(P, 4)
        dLt = torch.tensor(dL, device=dev)
        lib.ref_voxel_backward(P, R, nx, ny, nz, f(sV[0]), f(sV[1]), f(sV[2]), f(ctr[0]), f(ctr[1]), f(ctr[2]),
                               _ptr(t["means"]), _ptr(t["scales"]), f(1.0), _ptr(t["rots"]), None, _ptr(rx), _ptr(ry),
                               _ptr(rz), _ptr(dLt), _ptr(gn), _ptr(gc), _ptr(go), _ptr(g3), _ptr(gcov), _ptr(gs), _ptr(gr))
        torch.cuda.synchronize()
        res["grads"] = dict(dL_dopacity=go.cpu().numpy(), dL_dmean3D=g3.cpu().numpy(), dL_dcov3D=gcov.cpu().numpy(),
                            dL_dscale=gs.cpu().numpy(), dL_drot=gr.cpu().numpy())
    return res


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
