"""End-to-end check of the north star's "at matched 3D PSNR": the same reconstruction (same data, seeds, schedule,
densification rules) is trained twice --

  ours       r2_gaussian_b200 render()/query() + fused loss kernels + FusedAdam
  reference  the reference's own CUDA rasterizer / voxelizer (oracle/_ref, bridged to autograd here, bench-only)
             + the reference's loss formulas in torch ops + torch.optim.Adam

-- and 3-D PSNR (query() volume vs ground truth), 2-D PSNR on held-out views and time per iteration are reported.
Ground truth is a structured Gaussian-mixture phantom, so its projections / volume are exact for both arms.

    python scripts/train_compare.py [--iters 1500] [--det 256] [--arms ours,reference]
"""
import argparse
import ctypes as C
import json
import math
import os
import sys
import time
import types

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from r2_gaussian_b200 import losses, scene  # noqa: E402
from r2_gaussian_b200.gaussian_model import GaussianModel  # noqa: E402
from r2_gaussian_b200.render_query import query, render  # noqa: E402

DEV = "cuda"
PIPE = types.SimpleNamespace(compute_cov3D_python=False, debug=False)


# ------------------------------------------------------------------ phantom
def phantom(seed=0):
    rng = np.random.default_rng(seed)

    def ball(n, c, r, rho):
        d = rng.normal(size=(n, 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        p = c + d * (r * rng.uniform(0, 1, size=(n, 1)) ** (1 / 3))
        return p, np.full((n, 1), rho)

    parts = [ball(14000, np.array([0.0, 0.0, 0.0]), np.array([0.75, 0.6, 0.5]), 0.20),
             ball(5000, np.array([0.25, 0.1, 0.0]), np.array([0.25, 0.25, 0.25]), 0.55),
             ball(4000, np.array([-0.3, -0.1, 0.1]), np.array([0.2, 0.3, 0.2]), 0.40),
             ball(1500, np.array([0.0, 0.3, -0.2]), np.array([0.08, 0.08, 0.08]), 0.90)]
    xyz = np.concatenate([p for p, _ in parts]).astype(np.float32)
    rho = np.concatenate([r for _, r in parts]).astype(np.float32)
    d2 = np.maximum(scene.knn3_mean_sq_dist(xyz), 1e-6)
    s = np.clip(np.sqrt(d2) * 1.2, 0.01, 0.2).astype(np.float32)
    # densities scaled so that overlapping blobs add up to ~rho
    return scene.Cloud(xyz, np.repeat(s[:, None], 3, 1), np.tile(np.float32([1, 0, 0, 0]), (len(xyz), 1)),
                       (rho * 0.35).astype(np.float32))


class Fixed:
    """Duck-typed model of fixed (activated) parameters for render()/query()."""

    def __init__(self, cloud):
        t = lambda a: torch.tensor(a, device=DEV)
        self.get_xyz, self.get_density = t(cloud.means), t(cloud.density)
        self.get_scaling, self.get_rotation = t(cloud.scales), t(cloud.rotations)


# ------------------------------------------------------------------ reference arm (bench-only bridge)
def ref_lib():
    import util
    lib = util.ref_lib()
    if lib is None:
        raise RuntimeError("oracle/_ref/libr2ref.so is not built")
    lib.ref_raster_forward.restype = C.c_int
    lib.ref_voxel_forward.restype = C.c_int
    return lib


_vp = lambda t: C.c_void_p(t.data_ptr())
_f = C.c_float


class RefRaster(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means, means2D, dens, scales, rots, cam, lib):
        P, H, W = means.shape[0], cam.image_height, cam.image_width
        out = torch.zeros((1, H, W), device=DEV)
        radii = torch.zeros(P, dtype=torch.int32, device=DEV)
        tx, ty = (math.tan(cam.FoVx * 0.5), math.tan(cam.FoVy * 0.5)) if cam.mode == 1 else (1.0, 1.0)
        a = [t.contiguous() for t in (means, dens, scales, rots)]
        R = lib.ref_raster_forward(P, W, H, _vp(a[0]), _vp(a[1]), _vp(a[2]), _f(1.0), _vp(a[3]), None,
                                   _vp(cam.world_view_transform), _vp(cam.full_proj_transform), _vp(cam.camera_center),
                                   _f(tx), _f(ty), int(cam.mode), _vp(out), _vp(radii))
        ctx.save_for_backward(*a, radii)
        ctx.misc = (cam, lib, R, tx, ty)
        ctx.mark_non_differentiable(radii)
        return out, radii

    @staticmethod
    def backward(ctx, g, _):
        means, dens, scales, rots, radii = ctx.saved_tensors
        cam, lib, R, tx, ty = ctx.misc
        P, H, W = means.shape[0], cam.image_height, cam.image_width
        z = lambda *s: torch.zeros(s, device=DEV)
        g2, gc, go, gm, g3, gcov, gs, gr = z(P, 3), z(P, 4), z(P, 1), z(P, 1), z(P, 3), z(P, 6), z(P, 3), z(P, 4)
        lib.ref_raster_backward(P, R, W, H, _vp(means), _vp(scales), _f(1.0), _vp(rots), None,
                                _vp(cam.world_view_transform), _vp(cam.full_proj_transform), _vp(cam.camera_center),
                                _f(tx), _f(ty), _vp(radii), _vp(g.contiguous()), _vp(g2), _vp(gc), _vp(go), _vp(gm),
                                _vp(g3), _vp(gcov), _vp(gs), _vp(gr), int(cam.mode))
        return g3, g2, go, gs, gr, None, None


class RefVoxel(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means, dens, scales, rots, center, nV, sV, lib):
        P = means.shape[0]
        vol = torch.zeros(tuple(nV), device=DEV)
        rx = torch.zeros(P, dtype=torch.int32, device=DEV)
        ry, rz = torch.zeros_like(rx), torch.zeros_like(rx)
        a = [t.contiguous() for t in (means, dens, scales, rots)]
        R = lib.ref_voxel_forward(P, nV[0], nV[1], nV[2], _f(sV[0]), _f(sV[1]), _f(sV[2]), _f(center[0]), _f(center[1]),
                                  _f(center[2]), _vp(a[0]), _vp(a[1]), _vp(a[2]), _f(1.0), _vp(a[3]), None, _vp(vol),
                                  _vp(rx), _vp(ry), _vp(rz))
        ctx.save_for_backward(*a, rx, ry, rz)
        ctx.misc = (center, nV, sV, lib, R)
        return vol

    @staticmethod
    def backward(ctx, g):
        means, dens, scales, rots, rx, ry, rz = ctx.saved_tensors
        center, nV, sV, lib, R = ctx.misc
        P = means.shape[0]
        z = lambda *s: torch.zeros(s, device=DEV)
        gn, gc6, go, g3, gcov, gs, gr = z(P, 3), z(P, 6), z(P, 1), z(P, 3), z(P, 6), z(P, 3), z(P, 4)
        lib.ref_voxel_backward(P, R, nV[0], nV[1], nV[2], _f(sV[0]), _f(sV[1]), _f(sV[2]), _f(center[0]), _f(center[1]),
                               _f(center[2]), _vp(means), _vp(scales), _f(1.0), _vp(rots), None, _vp(rx), _vp(ry),
                               _vp(rz), _vp(g.contiguous()), _vp(gn), _vp(gc6), _vp(go), _vp(g3), _vp(gcov), _vp(gs),
                               _vp(gr))
        return g3, go, gs, gr, None, None, None, None


def ref_render(cam, pc, lib):
    xyz = pc.get_xyz
    vsp = torch.zeros_like(xyz, requires_grad=True) + 0
    if vsp.requires_grad:
        vsp.retain_grad()
    img, radii = RefRaster.apply(xyz, vsp, pc.get_density, pc.get_scaling, pc.get_rotation, cam, lib)
    return {"render": img, "viewspace_points": vsp, "visibility_filter": radii > 0, "radii": radii}


def ref_query(pc, center, nV, sV, lib):
    return {"vol": RefVoxel.apply(pc.get_xyz, pc.get_density, pc.get_scaling, pc.get_rotation, list(center), list(nV),
                                  list(sV), lib)}


def torch_window():
    g = torch.tensor([math.exp(-((x - 5) ** 2) / float(2 * 1.5 ** 2)) for x in range(11)])
    g = (g / g.sum()).unsqueeze(1)
    return g.mm(g.t()).float()[None, None].to(DEV)


def torch_losses(img, gt, w, lam):
    l1 = (img - gt).abs().mean()
    A, B = img[None], gt[None]
    mu1, mu2 = F.conv2d(A, w, padding=5), F.conv2d(B, w, padding=5)
    s11 = F.conv2d(A * A, w, padding=5) - mu1 * mu1
    s22 = F.conv2d(B * B, w, padding=5) - mu2 * mu2
    s12 = F.conv2d(A * B, w, padding=5) - mu1 * mu2
    m = ((2 * mu1 * mu2 + 1e-4) * (2 * s12 + 9e-4)) / ((mu1 * mu1 + mu2 * mu2 + 1e-4) * (s11 + s22 + 9e-4))
    return l1 + lam * (1 - m.mean())


def torch_tv(vol):
    nx, ny, nz = vol.shape
    tv = vol.diff(dim=0).abs().sum() + vol.diff(dim=1).abs().sum() + vol.diff(dim=2).abs().sum()
    return tv / ((nx - 1) * ny * nz + nx * (ny - 1) * nz + nx * ny * (nz - 1))


# ------------------------------------------------------------------ training
def psnr(a, b):
    mse = ((a - b) ** 2).mean().item()
    return 10 * math.log10(b.max().item() ** 2 / max(mse, 1e-20))


def train(arm, args, data):
    torch.manual_seed(0)
    np.random.seed(0)
    cams, gts, test_cams, test_gts, gt_vol, init_xyz, init_rho = data
    lib = ref_lib() if arm == "reference" else None
    gm = GaussianModel((0.0005 * 2, 0.5 * 2))
    gm.create_from_pcd(init_xyz, init_rho, 1.0)
    it_n = args.iters
    opt = types.SimpleNamespace(
        position_lr_init=2e-4, position_lr_final=2e-5, position_lr_max_steps=it_n,
        density_lr_init=1e-2, density_lr_final=1e-3, density_lr_max_steps=it_n,
        scaling_lr_init=5e-3, scaling_lr_final=5e-4, scaling_lr_max_steps=it_n,
        rotation_lr_init=1e-3, rotation_lr_final=1e-4, rotation_lr_max_steps=it_n)
    gm.training_setup(opt)
    if arm == "reference":   # the reference's optimizer
        gm.optimizer = torch.optim.Adam([{"params": g["params"], "lr": g["lr"], "name": g["name"]}
                                         for g in gm.optimizer.param_groups], lr=0.0, eps=1e-15)
    w = torch_window()
    bbox = torch.tensor([[-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]], device=DEV)
    tv_n, tv_s = [32, 32, 32], [0.25, 0.25, 0.25]
    rng = np.random.default_rng(1)
    order = rng.integers(0, len(cams), size=it_n + 1)
    centers = (-1 + 0.125) + (2 - 0.25) * rng.random((it_n + 1, 3))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for it in range(1, it_n + 1):
        gm.update_learning_rate(it)
        cam, gt = cams[order[it]], gts[order[it]]
        if arm == "ours":
            pkg = render(cam, gm, PIPE)
            loss = losses.image_loss(pkg["render"], gt, 0.25)["total"]
            vol = query(gm, centers[it], tv_n, tv_s, PIPE)["vol"]
            loss = loss + 0.05 * losses.tv_3d_loss(vol, "mean")
        else:
            pkg = ref_render(cam, gm, lib)
            loss = torch_losses(pkg["render"], gt, w, 0.25)
            vol = ref_query(gm, centers[it], tv_n, tv_s, lib)["vol"]
            loss = loss + 0.05 * torch_tv(vol)
        loss.backward()
        with torch.no_grad():
            vis = pkg["visibility_filter"]
            gm.update_max_radii(pkg["radii"], vis)
            gm.add_densification_stats(pkg["viewspace_points"], vis)
            if args.densify_from < it < args.densify_until and it % 100 == 0:
                gm.densify_and_prune(args.grad_thr, 1e-5, None, None, 500000, 0.1 * 2.0 * 0.1, bbox)
            gm.optimizer.step()
            gm.optimizer.zero_grad(set_to_none=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    with torch.no_grad():
        if arm == "ours":
            vol = query(gm, [0, 0, 0], [128] * 3, [2.0] * 3, PIPE)["vol"]
            p2 = np.mean([psnr(render(c, gm, PIPE)["render"], g) for c, g in zip(test_cams, test_gts)])
        else:
            vol = ref_query(gm, [0, 0, 0], [128] * 3, [2.0] * 3, lib)["vol"]
            p2 = np.mean([psnr(ref_render(c, gm, lib)["render"], g) for c, g in zip(test_cams, test_gts)])
    return {"psnr_3d": psnr(vol, gt_vol), "psnr_2d_test": float(p2), "gaussians": int(gm.get_xyz.shape[0]),
            "seconds": dt, "ms_per_iteration": dt / it_n * 1e3, "final_loss": float(loss)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=1500)
    ap.add_argument("--det", type=int, default=256)
    ap.add_argument("--views", type=int, default=50)
    ap.add_argument("--init-points", type=int, default=5000)
    ap.add_argument("--densify-from", type=int, default=300)
    ap.add_argument("--densify-until", type=int, default=1200)
    ap.add_argument("--grad-thr", type=float, default=5e-5)
    ap.add_argument("--arms", default="ours,reference")
    args = ap.parse_args()
    truth = Fixed(phantom())
    scanner = scene.cone_beam_scanner(args.det)
    allv = [scene.camera_from_view(v) for v in scene.make_views(scanner, args.views + 10)]
    test_idx = set(range(2, len(allv), (len(allv)) // 10)[:10])
    cams = [c for i, c in enumerate(allv) if i not in test_idx]
    test_cams = [c for i, c in enumerate(allv) if i in test_idx]
    with torch.no_grad():
        gts = [render(c, truth, PIPE)["render"].clone() for c in cams]
        test_gts = [render(c, truth, PIPE)["render"].clone() for c in test_cams]
        gt_vol = query(truth, [0, 0, 0], [128] * 3, [2.0] * 3, PIPE)["vol"].clone()
        coarse = query(truth, [0, 0, 0], [64] * 3, [2.0] * 3, PIPE)["vol"]
    # initial cloud: random voxels of a noisy coarse volume above 5 % of its maximum (initialize_pcd.py's recipe)
    g = torch.Generator(DEV).manual_seed(0)
    noisy = (coarse + 0.05 * coarse.max() * torch.randn(coarse.shape, device=DEV, generator=g)).clamp_min(0)
    idx = torch.nonzero(noisy > 0.05 * noisy.max())
    pick = idx[torch.randperm(idx.shape[0], device=DEV, generator=g)[: args.init_points]]
    init_xyz = ((pick.float() + 0.5) / 64 * 2 - 1).cpu().numpy().astype(np.float32)
    init_rho = (noisy[pick[:, 0], pick[:, 1], pick[:, 2]] * 0.15).clamp_min(1e-3)[:, None].cpu().numpy().astype(np.float32)
    data = (cams, gts, test_cams, test_gts, gt_vol, init_xyz, init_rho)
    out = {"config": {"iters": args.iters, "detector": args.det, "train_views": len(cams), "test_views": len(test_cams),
                      "init_points": int(init_xyz.shape[0]), "truth_gaussians": int(truth.get_xyz.shape[0])}}
    for arm in args.arms.split(","):
        out[arm] = train(arm, args, data)
    if "ours" in out and "reference" in out:
        out["speedup_per_iteration"] = out["reference"]["ms_per_iteration"] / out["ours"]["ms_per_iteration"]
        out["psnr_3d_delta_db"] = out["ours"]["psnr_3d"] - out["reference"]["psnr_3d"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
