"""Secondary measurements (not the driver's headline): training-style forward+backward of the projector,
the voxelizer sweep (256^3 over 500k Gaussians) and the 32^3 TV crop, ours vs the compiled reference
(oracle/_ref).  Prints one JSON object.  Same protocol as bench.py: CUDA events per step, L2 flushed
between steps."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from r2_gaussian_b200 import _C, scene  # noqa: E402

dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
sync = torch.cuda.synchronize
E = torch.Tensor([])
out = {}


def timed(fn, steps=30, warmup=5):
    return float(np.mean(bench.timed_steps(lambda i: fn(i), steps, warmup, flush, sync)))


class A:
    gaussians = 100000; detector = 512; views = 50; cloud = "init"


sc, views, cloud = bench.build_scene(A)
m = torch.tensor(cloud.means, device=dev); s = torch.tensor(cloud.scales, device=dev)
r = torch.tensor(cloud.rotations, device=dev); d = torch.tensor(cloud.density, device=dev)
dv = bench.device_views(views, dev)
dL = torch.randn(1, 512, 512, device=dev)

# ---- projector forward + backward through the extension-level API (the reference _C signature) ----
def ours_fb(i):
    v = dv[i % 50]
    R, img, radii, geom, binning, imgb = _C.rasterize_gaussians(m, d, s, r, 1.0, E, v["view"], v["proj"], v["tx"], v["ty"], 512, 512, v["campos"], False, v["mode"], False)
    _C.rasterize_gaussians_backward(m, radii, s, r, 1.0, E, v["view"], v["proj"], v["tx"], v["ty"], dL, v["campos"], geom, R, binning, imgb, v["mode"], False)

out["raster_fwd_bwd_ms_ours"] = timed(ours_fb)

state = {}
def ours_f(i):
    v = dv[i % 50]
    state["x"] = _C.rasterize_gaussians(m, d, s, r, 1.0, E, v["view"], v["proj"], v["tx"], v["ty"], 512, 512, v["campos"], False, v["mode"], False)
out["raster_fwd_ms_ours_sync_api"] = timed(ours_f)
def ours_b(i):
    v = dv[0]
    R, img, radii, geom, binning, imgb = state["x"]
    _C.rasterize_gaussians_backward(m, radii, s, r, 1.0, E, v["view"], v["proj"], v["tx"], v["ty"], dL, v["campos"], geom, R, binning, imgb, v["mode"], False)
v0 = dv[0]
state["x"] = _C.rasterize_gaussians(m, d, s, r, 1.0, E, v0["view"], v0["proj"], v0["tx"], v0["ty"], 512, 512, v0["campos"], False, v0["mode"], False)
out["raster_bwd_ms_ours"] = timed(ours_b)

ref_path = os.path.join(ROOT, "oracle", "_ref", "libr2ref.so")
lib = C.CDLL(ref_path) if os.path.exists(ref_path) else None
vp = lambda t: C.c_void_p(t.data_ptr())
f = C.c_float
if lib is not None:
    lib.ref_raster_forward.restype = C.c_int
    lib.ref_voxel_forward.restype = C.c_int
    P = cloud.P
    o = torch.zeros((1, 512, 512), device=dev); radii = torch.zeros(P, dtype=torch.int32, device=dev)
    z = lambda *sh: torch.zeros(sh, device=dev)
    g2, gc, go, gm, g3, gcov, gs, gr = z(P, 3), z(P, 4), z(P, 1), z(P, 1), z(P, 3), z(P, 6), z(P, 3), z(P, 4)

    def ref_fb(i):
        v = dv[i % 50]
        o.zero_(); radii.zero_()
        R = lib.ref_raster_forward(P, 512, 512, vp(m), vp(d), vp(s), f(1.0), vp(r), None, vp(v["view"]), vp(v["proj"]), vp(v["campos"]), f(v["tx"]), f(v["ty"]), int(v["mode"]), vp(o), vp(radii))
        for t in (g2, gc, go, gm, g3, gcov, gs, gr):
            t.zero_()
        lib.ref_raster_backward(P, R, 512, 512, vp(m), vp(s), f(1.0), vp(r), None, vp(v["view"]), vp(v["proj"]), vp(v["campos"]), f(v["tx"]), f(v["ty"]), vp(radii), vp(dL), vp(g2), vp(gc), vp(go), vp(gm), vp(g3), vp(gcov), vp(gs), vp(gr), int(v["mode"]))
    out["raster_fwd_bwd_ms_ref"] = timed(ref_fb)

# ---- voxelizer sweep: 256^3 over 500k Gaussians (BASELINE config 5) ----
big = scene.make_cloud(500000, kind="init", seed=0)
bm = torch.tensor(big.means, device=dev); bs = torch.tensor(big.scales, device=dev)
br = torch.tensor(big.rotations, device=dev); bd = torch.tensor(big.density, device=dev)
from r2_gaussian_b200.engine import VoxelEngine
ve = VoxelEngine(big.P, (256, 256, 256), dev, capacity=14_000_000)
Rv = ve.fit(bm, bd, bs, br, (2.0, 2.0, 2.0), (0.0, 0.0, 0.0))
out["voxel256_R"] = Rv
out["voxel256_500k_fwd_ms_ours"] = timed(lambda i: ve.forward(bm, bd, bs, br, (2.0, 2.0, 2.0), (0.0, 0.0, 0.0)), 10, 2)
out["voxel256_render_only_ms_ours"] = timed(lambda i: ve.render_only(), 10, 2)
if lib is not None:
    vol = torch.zeros((256, 256, 256), device=dev)
    rx = torch.zeros(big.P, dtype=torch.int32, device=dev); ry = torch.zeros_like(rx); rz = torch.zeros_like(rx)
    def ref_v(i):
        vol.zero_()
        lib.ref_voxel_forward(big.P, 256, 256, 256, f(2.0), f(2.0), f(2.0), f(0.0), f(0.0), f(0.0), vp(bm), vp(bd), vp(bs), f(1.0), vp(br), None, vp(vol), vp(rx), vp(ry), vp(rz))
    out["voxel256_500k_fwd_ms_ref"] = timed(ref_v, 5, 1)
    err = (ve.forward(bm, bd, bs, br, (2.0, 2.0, 2.0), (0.0, 0.0, 0.0)) - vol).abs().max().item() / vol.abs().max().item()
    out["voxel256_rel_err_vs_ref"] = err

# ---- TV crop: 32^3 sub-volume of the 100k cloud, forward + backward (train.py:128-144) ----
dV = torch.randn(32, 32, 32, device=dev)
def ours_tv(i):
    R, vol_, rx_, ry_, rz_, geom, binning, imgb = _C.voxelize_gaussians(m, d, s, r, 1.0, E, 32, 32, 32, 0.25, 0.25, 0.25, 0.3, -0.4, 0.1, False, False)
    _C.voxelize_gaussians_backward(m, rx_, ry_, rz_, s, r, 1.0, E, dV, geom, R, binning, imgb, 32, 32, 32, 0.25, 0.25, 0.25, 0.3, -0.4, 0.1, False)
out["tvcrop_fwd_bwd_ms_ours"] = timed(ours_tv)
if lib is not None:
    P = cloud.P
    vol2 = torch.zeros((32, 32, 32), device=dev)
    qx = torch.zeros(P, dtype=torch.int32, device=dev); qy = torch.zeros_like(qx); qz = torch.zeros_like(qx)
    gn, gc6, go1, g31, gcv, gs1, gr1 = z(P, 3), z(P, 6), z(P, 1), z(P, 3), z(P, 6), z(P, 3), z(P, 4)
    def ref_tv(i):
        vol2.zero_()
        R = lib.ref_voxel_forward(P, 32, 32, 32, f(0.25), f(0.25), f(0.25), f(0.3), f(-0.4), f(0.1), vp(m), vp(d), vp(s), f(1.0), vp(r), None, vp(vol2), vp(qx), vp(qy), vp(qz))
        for t in (gn, gc6, go1, g31, gcv, gs1, gr1):
            t.zero_()
        lib.ref_voxel_backward(P, R, 32, 32, 32, f(0.25), f(0.25), f(0.25), f(0.3), f(-0.4), f(0.1), vp(m), vp(s), f(1.0), vp(r), None, vp(qx), vp(qy), vp(qz), vp(dV), vp(gn), vp(gc6), vp(go1), vp(g31), vp(gcv), vp(gs1), vp(gr1))
    out["tvcrop_fwd_bwd_ms_ref"] = timed(ref_tv)
print(json.dumps(out))
