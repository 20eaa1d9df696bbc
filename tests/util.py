"""Shared helpers for the parity tests: scene cases, running our C ABI, the oracle and (GPU box only)
the compiled reference in oracle/_ref."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from r2_gaussian_b200 import scene

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libr2ref.so")


def case(name: str):
    """Named small scenes: returns (cloud, view)."""
    kind, beam, P, n = {
        "cone_init_small": ("init", "cone", 3000, 128),
        "cone_trained_small": ("trained", "cone", 3000, 128),
        "parallel_trained_small": ("trained", "parallel", 2000, 96),
        "cone_trained_ragged": ("trained", "cone", 1500, 100),   # detector not a multiple of 16
        "cone_trained_mid": ("trained", "cone", 20000, 256),
        "cone_init_mid": ("init", "cone", 50000, 256),
        "cone_trained_bigdet": ("trained", "cone", 1500, 1040),  # 65 x 65 = 4225 tiles > DIRECT_MAX_TILES: radix path
    }[name]
    sc = scene.cone_beam_scanner(n, 64) if beam == "cone" else scene.parallel_beam_scanner(n, 64)
    view = scene.make_view(sc, 0.37 + 0.1 * len(name))
    cloud = scene.make_cloud(P, kind=kind, seed=len(name))
    return cloud, view


def to_torch(cloud, view, device="cuda", requires_grad=False):
    import torch

    t = dict(
        means=torch.tensor(cloud.means, device=device), scales=torch.tensor(cloud.scales, device=device),
        rots=torch.tensor(cloud.rotations, device=device), dens=torch.tensor(cloud.density, device=device),
    )
    if requires_grad:
        for v in t.values():
            v.requires_grad_(True)
    if view is not None:
        t["view"] = torch.tensor(view.viewmatrix, device=device)
        t["proj"] = torch.tensor(view.projmatrix, device=device)
        t["campos"] = torch.tensor(view.campos, device=device)
    return t


def ours_raster_forward(cloud, view, export=True, cov3D_precomp=None, debug=False):
    import torch
    from r2_gaussian_b200 import _C
    from r2_gaussian_b200._lib import load, check

    t = to_torch(cloud, view)
    empty = torch.Tensor([])
    scales, rots, cov = t["scales"], t["rots"], empty
    if cov3D_precomp is not None:
        scales, rots, cov = empty, empty, torch.tensor(cov3D_precomp, device="cuda")
    R, color, radii, geom, binning, img = _C.rasterize_gaussians(
        t["means"], t["dens"], scales, rots, 1.0, cov, t["view"], t["proj"], view.tanfovx, view.tanfovy,
        view.image_height, view.image_width, t["campos"], False, view.mode, debug)
    t["scales_in"], t["rots_in"], t["cov_in"] = scales, rots, cov
    out = dict(R=R, image=color[0].cpu().numpy(), radii=radii.cpu().numpy(), state=(geom, binning, img), t=t)
    if export:
        lib = load()
        P, W, H = cloud.P, view.image_width, view.image_height
        T = ((W + 15) // 16) * ((H + 15) // 16)
        dev = "cuda"
        xy = torch.empty((P, 2), device=dev); depth = torch.empty(P, device=dev)
        co = torch.empty((P, 4), device=dev); mu = torch.empty(P, device=dev)
        tt = torch.empty(P, dtype=torch.int32, device=dev); po = torch.empty(P, dtype=torch.int32, device=dev)
        keys = torch.empty(max(R, 1), dtype=torch.int64, device=dev)
        pl = torch.empty(max(R, 1), dtype=torch.int32, device=dev)
        ranges = torch.empty((T, 2), dtype=torch.int32, device=dev)
        from r2_gaussian_b200._C import _carved_capacity
        cap = _carved_capacity(binning, R)
        keys = torch.empty(max(cap, 1), dtype=torch.int64, device=dev)
        pl = torch.empty(max(cap, 1), dtype=torch.int32, device=dev)
        rc = lib.r2x_raster_export(torch.cuda.current_stream().cuda_stream, P, W, H, cap, geom.data_ptr(),
                                   binning.data_ptr() if binning.numel() else None, img.data_ptr(), xy.data_ptr(),
                                   depth.data_ptr(), co.data_ptr(), mu.data_ptr(), tt.data_ptr(), po.data_ptr(),
                                   keys.data_ptr(), pl.data_ptr(), ranges.data_ptr())
        check(rc, "r2x_raster_export")
        torch.cuda.synchronize()
        out.update(xy=xy.cpu().numpy(), depth=depth.cpu().numpy(), conic_opacity=co.cpu().numpy(), mu=mu.cpu().numpy(),
                   tiles_touched=tt.cpu().numpy().astype(np.uint32), point_offsets=po.cpu().numpy().astype(np.uint32),
                   keys=keys.cpu().numpy().astype(np.uint64)[:R], point_list=pl.cpu().numpy().astype(np.uint32)[:R],
                   ranges=ranges.cpu().numpy().astype(np.uint32))
    return out


def ours_raster_backward(cloud, view, fwd, dL, debug=False):
    import torch
    from r2_gaussian_b200 import _C

    t = fwd["t"]
    geom, binning, img = fwd["state"]
    radii = torch.tensor(fwd["radii"], device="cuda")
    g = _C.rasterize_gaussians_backward(
        t["means"], radii, t["scales_in"], t["rots_in"], 1.0, t["cov_in"], t["view"], t["proj"], view.tanfovx,
        view.tanfovy, torch.tensor(dL, device="cuda")[None], t["campos"], geom, fwd["R"], binning, img, view.mode, debug)
    names = ["dL_dmean2D", "dL_dopacity", "dL_dmu", "dL_dmean3D", "dL_dcov3D", "dL_dscale", "dL_drot"]
    return {n: x.cpu().numpy() for n, x in zip(names, g)}


def oracle_raster_forward(cloud, view, **kw):
    from oracle import r2_oracle as orc

    return orc.raster_forward(cloud.means, cloud.scales, cloud.rotations, cloud.density, view.viewmatrix,
                              view.projmatrix, view.image_width, view.image_height, view.tanfovx, view.tanfovy,
                              view.mode, **kw)


def oracle_raster_backward(cloud, view, fwd, dL):
    from oracle import r2_oracle as orc

    return orc.raster_backward(fwd, cloud.means, cloud.scales, cloud.rotations, view.viewmatrix, view.projmatrix,
                               view.image_width, view.image_height, view.tanfovx, view.tanfovy, view.mode, dL)


# ---- voxelizer ------------------------------------------------------------------------------
def ours_voxel_forward(cloud, nVoxel, sVoxel, center, export=True, cov3D_precomp=None, debug=False):
    import torch
    from r2_gaussian_b200 import _C
    from r2_gaussian_b200._lib import load, check

    t = to_torch(cloud, None)
    rots, cov = t["rots"], torch.Tensor([])
    if cov3D_precomp is not None:   # the voxelizer needs the scales for its bounding radius either way
        rots, cov = torch.Tensor([]), torch.tensor(cov3D_precomp, device="cuda")
    t["rots_in"], t["cov_in"] = rots, cov
    R, vol, rx, ry, rz, geom, binning, img = _C.voxelize_gaussians(
        t["means"], t["dens"], t["scales"], rots, 1.0, cov, nVoxel[0], nVoxel[1], nVoxel[2],
        sVoxel[0], sVoxel[1], sVoxel[2], center[0], center[1], center[2], False, debug)
    out = dict(R=R, vol=vol.cpu().numpy(), radii_x=rx.cpu().numpy(), radii_y=ry.cpu().numpy(), radii_z=rz.cpu().numpy(),
               state=(geom, binning, img), t=t, radii_t=(rx, ry, rz))
    if export:
        lib = load()
        P = cloud.P
        nx, ny, nz = nVoxel
        T = ((nx + 7) // 8) * ((ny + 7) // 8) * ((nz + 7) // 8)
        dev = "cuda"
        xyz = torch.empty((P, 3), device=dev); depth = torch.empty(P, device=dev); co = torch.empty((P, 7), device=dev)
        tt = torch.empty(P, dtype=torch.int32, device=dev); po = torch.empty(P, dtype=torch.int32, device=dev)
        keys = torch.empty(max(R, 1), dtype=torch.int64, device=dev)
        pl = torch.empty(max(R, 1), dtype=torch.int32, device=dev)
        ranges = torch.empty((T, 2), dtype=torch.int32, device=dev)
        from r2_gaussian_b200._C import _carved_capacity
        cap = _carved_capacity(binning, R)
        keys = torch.empty(max(cap, 1), dtype=torch.int64, device=dev)
        pl = torch.empty(max(cap, 1), dtype=torch.int32, device=dev)
        rc = lib.r2x_voxel_export(torch.cuda.current_stream().cuda_stream, P, nx, ny, nz, cap, geom.data_ptr(),
                                  binning.data_ptr() if binning.numel() else None, img.data_ptr(), xyz.data_ptr(),
                                  depth.data_ptr(), co.data_ptr(), tt.data_ptr(), po.data_ptr(), keys.data_ptr(),
                                  pl.data_ptr(), ranges.data_ptr())
        check(rc, "r2x_voxel_export")
        torch.cuda.synchronize()
        out.update(xyz_vol=xyz.cpu().numpy(), depth=depth.cpu().numpy(), conic_opacity=co.cpu().numpy(),
                   tiles_touched=tt.cpu().numpy().astype(np.uint32), keys=keys.cpu().numpy().astype(np.uint64)[:R],
                   point_list=pl.cpu().numpy().astype(np.uint32)[:R], ranges=ranges.cpu().numpy().astype(np.uint32))
    return out


def ours_voxel_backward(cloud, nVoxel, sVoxel, center, fwd, dL, debug=False):
    import torch
    from r2_gaussian_b200 import _C

    t = fwd["t"]
    geom, binning, img = fwd["state"]
    rx, ry, rz = fwd["radii_t"]
    g = _C.voxelize_gaussians_backward(
        t["means"], rx, ry, rz, t["scales"], t["rots_in"], 1.0, t["cov_in"], torch.tensor(dL, device="cuda"), geom,
        fwd["R"], binning, img, nVoxel[0], nVoxel[1], nVoxel[2], sVoxel[0], sVoxel[1], sVoxel[2], center[0], center[1],
        center[2], debug)
    names = ["dL_dopacity", "dL_dmean3D", "dL_dcov3D", "dL_dscale", "dL_drot"]
    return {n: x.cpu().numpy() for n, x in zip(names, g)}


def oracle_voxel_forward(cloud, nVoxel, sVoxel, center, **kw):
    from oracle import r2_oracle as orc

    return orc.voxel_forward(cloud.means, cloud.scales, cloud.rotations, cloud.density, nVoxel, sVoxel, center, **kw)


def oracle_voxel_backward(cloud, nVoxel, sVoxel, fwd, dL):
    from oracle import r2_oracle as orc

    return orc.voxel_backward(fwd, cloud.scales, cloud.rotations, nVoxel, sVoxel, dL)


# ---- comparison helpers ---------------------------------------------------------------------
def key_multiset_equal(keys_a, keys_b):
    return np.array_equal(np.sort(np.asarray(keys_a, dtype=np.uint64)), np.sort(np.asarray(keys_b, dtype=np.uint64)))


def rel_err(a, b, floor=None):
    """max |a-b| / max(|b|, floor) -- floor defaults to 1e-3 * max|b| (relative to the signal scale)."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    if floor is None:
        floor = 1e-3 * (np.abs(b).max() if b.size else 1.0) + 1e-30
    return float((np.abs(a - b) / np.maximum(np.abs(b), floor)).max()) if a.size else 0.0


def grad_mismatch(a, b, rtol=2e-4, atol_rel=2e-5, scale=None):
    """Worst violation of |a-b| <= rtol*|b| + atol_rel*scale (scale defaults to max|b|); <= 1 passes.

    Gradients are sums of thousands of signed float32 terms: the achievable agreement between two
    summation orders (the reference's float atomics are not even run-to-run stable) is relative to the
    magnitude of the terms, not of the (possibly cancelling) sum, hence the absolute part tied to the
    array's scale."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    if scale is None:
        scale = float(np.abs(b).max())
    tol = rtol * np.abs(b) + atol_rel * scale + 1e-30
    return float((np.abs(a - b) / tol).max())


def assert_grads_close(got: dict, want: dict, keys, rtol=2e-4, atol_rel=2e-5, label=""):
    for k in keys:
        scale = None
        if k == "dL_drot":
            # for isotropic Gaussians the rotation gradient is pure cancellation noise: tie its scale
            # to the scale gradient (same chain, dL/dM times a parameter of order one)
            scale = max(float(np.abs(want[k]).max()), float(np.abs(want["dL_dscale"]).max()))
        m = grad_mismatch(got[k], want[k], rtol, atol_rel, scale)
        assert m <= 1.0, f"{label}{k}: mismatch {m:.3g} x tolerance"


# ---- the compiled reference (GPU box) ---------------------------------------------------------
_ref = None


def ref_lib():
    global _ref
    if _ref is None:
        if not os.path.exists(REF_LIB):
            return None
        _ref = C.CDLL(REF_LIB)
        _ref.ref_raster_forward.restype = C.c_int
        _ref.ref_voxel_forward.restype = C.c_int
    return _ref


def need_ref():
    """The compiled reference (oracle/_ref/libr2ref.so).  It is built in the build container (where
    /root/reference exists) and travels to the GPU box; a GPU box without it is a broken snapshot, so
    the parity tests FAIL there instead of skipping."""
    import pytest

    lib = ref_lib()
    if lib is None:
        pytest.fail("oracle/_ref/libr2ref.so is missing: run `bash oracle/build_ref.sh` in the build container "
                    "(needs /root/reference) before sending the tree to the GPU box")
    return lib


def _cptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None and t.numel() else None


def run_ref_raster(cloud, view, dL=None, cov3D_precomp=None):
    import torch

    lib = need_ref()
    t = to_torch(cloud, view)
    P, W, H = cloud.P, view.image_width, view.image_height
    dev = "cuda"
    out = torch.zeros((1, H, W), device=dev); radii = torch.zeros(P, dtype=torch.int32, device=dev)
    f = C.c_float
    covp = None if cov3D_precomp is None else torch.tensor(cov3D_precomp, device=dev)
    sc_p = None if covp is not None else _cptr(t["scales"])
    ro_p = None if covp is not None else _cptr(t["rots"])
    R = lib.ref_raster_forward(P, W, H, _cptr(t["means"]), _cptr(t["dens"]), sc_p, f(1.0), ro_p,
                               _cptr(covp), _cptr(t["view"]), _cptr(t["proj"]), _cptr(t["campos"]), f(view.tanfovx),
                               f(view.tanfovy), view.mode, _cptr(out), _cptr(radii))
    torch.cuda.synchronize()
    T = ((W + 15) // 16) * ((H + 15) // 16)
    depth = torch.zeros(P, device=dev); xy = torch.zeros((P, 2), device=dev); cov = torch.zeros((P, 6), device=dev)
    co = torch.zeros((P, 4), device=dev); mu = torch.zeros(P, device=dev)
    tt = torch.zeros(P, dtype=torch.int32, device=dev); po = torch.zeros(P, dtype=torch.int32, device=dev)
    ks = torch.zeros(max(R, 1), dtype=torch.int64, device=dev); pl = torch.zeros(max(R, 1), dtype=torch.int32, device=dev)
    rg = torch.zeros((T, 2), dtype=torch.int32, device=dev); nc = torch.zeros((H, W), dtype=torch.int32, device=dev)
    lib.ref_raster_export(P, W, H, R, _cptr(depth), _cptr(xy), _cptr(cov), _cptr(co), _cptr(mu), _cptr(tt), _cptr(po), None,
                          None, _cptr(ks), _cptr(pl), _cptr(rg), _cptr(nc))
    res = dict(R=R, image=out[0].cpu().numpy(), radii=radii.cpu().numpy(), depth=depth.cpu().numpy(),
               xy=xy.cpu().numpy(), cov3D=cov.cpu().numpy(), conic_opacity=co.cpu().numpy(), mu=mu.cpu().numpy(),
               tiles_touched=tt.cpu().numpy().astype(np.uint32), keys=ks.cpu().numpy().astype(np.uint64)[:R],
               point_list=pl.cpu().numpy().astype(np.uint32)[:R], ranges=rg.cpu().numpy().astype(np.uint32))
    if dL is not None:
        z = lambda *s: torch.zeros(s, device=dev)
        g2, gc, go, gm, g3, gcov, gs, gr = z(P, 3), z(P, 4), z(P, 1), z(P, 1), z(P, 3), z(P, 6), z(P, 3), z(P, 4)
        dLt = torch.tensor(dL, device=dev)
        lib.ref_raster_backward(P, R, W, H, _cptr(t["means"]), sc_p, f(1.0), ro_p, _cptr(covp),
                                _cptr(t["view"]), _cptr(t["proj"]), _cptr(t["campos"]), f(view.tanfovx), f(view.tanfovy),
                                _cptr(radii), _cptr(dLt), _cptr(g2), _cptr(gc), _cptr(go), _cptr(gm), _cptr(g3), _cptr(gcov),
                                _cptr(gs), _cptr(gr), view.mode)
        torch.cuda.synchronize()
        res["grads"] = dict(dL_dmean2D=g2.cpu().numpy(), dL_dopacity=go.cpu().numpy(), dL_dmu=gm.cpu().numpy(),
                            dL_dmean3D=g3.cpu().numpy(), dL_dcov3D=gcov.cpu().numpy(), dL_dscale=gs.cpu().numpy(),
                            dL_drot=gr.cpu().numpy())
    return res


def run_ref_voxel(cloud, nV, sV, ctr, dL=None, cov3D_precomp=None):
    import torch

    lib = need_ref()
    t = to_torch(cloud, None)
    P = cloud.P
    nx, ny, nz = nV
    dev = "cuda"
    f = C.c_float
    vol = torch.zeros(nV, device=dev)
    covp = None if cov3D_precomp is None else torch.tensor(cov3D_precomp, device=dev)
    rx = torch.zeros(P, dtype=torch.int32, device=dev); ry = torch.zeros_like(rx); rz = torch.zeros_like(rx)
    R = lib.ref_voxel_forward(P, nx, ny, nz, f(sV[0]), f(sV[1]), f(sV[2]), f(ctr[0]), f(ctr[1]), f(ctr[2]),
                              _cptr(t["means"]), _cptr(t["dens"]), _cptr(t["scales"]), f(1.0),
                              None if covp is not None else _cptr(t["rots"]), _cptr(covp),
                              _cptr(vol), _cptr(rx), _cptr(ry), _cptr(rz))
    torch.cuda.synchronize()
    T = ((nx + 7) // 8) * ((ny + 7) // 8) * ((nz + 7) // 8)
    depth = torch.zeros(P, device=dev); xyz = torch.zeros((P, 3), device=dev); cov = torch.zeros((P, 6), device=dev)
    co = torch.zeros((P, 7), device=dev)
    tt = torch.zeros(P, dtype=torch.int32, device=dev); po = torch.zeros(P, dtype=torch.int32, device=dev)
    ks = torch.zeros(max(R, 1), dtype=torch.int64, device=dev); pl = torch.zeros(max(R, 1), dtype=torch.int32, device=dev)
    rg = torch.zeros((T, 2), dtype=torch.int32, device=dev)
    lib.ref_voxel_export(P, nx, ny, nz, R, _cptr(depth), _cptr(xyz), _cptr(cov), _cptr(co), _cptr(tt), _cptr(po), None, None,
                         _cptr(ks), _cptr(pl), _cptr(rg), None)
    res = dict(R=R, vol=vol.cpu().numpy(), radii_x=rx.cpu().numpy(), radii_y=ry.cpu().numpy(), radii_z=rz.cpu().numpy(),
               depth=depth.cpu().numpy(), xyz_vol=xyz.cpu().numpy(), conic_opacity=co.cpu().numpy(),
               tiles_touched=tt.cpu().numpy().astype(np.uint32), keys=ks.cpu().numpy().astype(np.uint64)[:R],
               point_list=pl.cpu().numpy().astype(np.uint32)[:R], ranges=rg.cpu().numpy().astype(np.uint32))
    if dL is not None:
        z = lambda *s: torch.zeros(s, device=dev)
        gn, gc, go, g3, gcov, gs, gr = z(P, 3), z(P, 6), z(P, 1), z(P, 3), z(P, 6), z(P, 3), z(P, 4)
        dLt = torch.tensor(dL, device=dev)
        lib.ref_voxel_backward(P, R, nx, ny, nz, f(sV[0]), f(sV[1]), f(sV[2]), f(ctr[0]), f(ctr[1]), f(ctr[2]),
                               _cptr(t["means"]), _cptr(t["scales"]), f(1.0),
                               None if covp is not None else _cptr(t["rots"]), _cptr(covp), _cptr(rx), _cptr(ry),
                               _cptr(rz), _cptr(dLt), _cptr(gn), _cptr(gc), _cptr(go), _cptr(g3), _cptr(gcov), _cptr(gs), _cptr(gr))
        torch.cuda.synchronize()
        res["grads"] = dict(dL_dopacity=go.cpu().numpy(), dL_dmean3D=g3.cpu().numpy(), dL_dcov3D=gcov.cpu().numpy(),
                            dL_dscale=gs.cpu().numpy(), dL_drot=gr.cpu().numpy())
    return res
