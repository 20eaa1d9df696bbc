"""Timings of the training-step pieces around the hot path against the reference's own
formulation in plain torch ops: image loss fwd+bwd, TV fwd+bwd, Adam step, 3-NN initialisation, densify+prune, and
one full training iteration (render + loss + backward + Adam).  Prints one JSON object."""
import json
import math
import sys
import os
import types

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from r2_gaussian_b200 import losses, scene  # noqa: E402
from r2_gaussian_b200.gaussian_model import GaussianModel  # noqa: E402
from r2_gaussian_b200.optim import FusedAdam  # noqa: E402
from r2_gaussian_b200.render_query import query, render  # noqa: E402
from r2_gaussian_b200.simple_knn import distCUDA2  # noqa: E402


def timeit(fn, n=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def torch_window(dev):
    g = torch.tensor([math.exp(-((x - 5) ** 2) / float(2 * 1.5 ** 2)) for x in range(11)])
    g = (g / g.sum()).unsqueeze(1)
    return g.mm(g.t()).float()[None, None].to(dev)


def torch_loss(a, b, w, lam):
    l1 = (a - b).abs().mean()
    A, B = a[None], b[None]
    mu1, mu2 = F.conv2d(A, w, padding=5), F.conv2d(B, w, padding=5)
    s11 = F.conv2d(A * A, w, padding=5) - mu1 * mu1
    s22 = F.conv2d(B * B, w, padding=5) - mu2 * mu2
    s12 = F.conv2d(A * B, w, padding=5) - mu1 * mu2
    m = ((2 * mu1 * mu2 + 1e-4) * (2 * s12 + 9e-4)) / ((mu1 * mu1 + mu2 * mu2 + 1e-4) * (s11 + s22 + 9e-4))
    return l1 + lam * (1 - m.mean())


def main():
    dev = "cuda"
    out = {}
    torch.manual_seed(0)
    a = torch.rand(1, 512, 512, device=dev, requires_grad=True)
    b = torch.rand(1, 512, 512, device=dev)
    w = torch_window(dev)

    def ref_loss():
        a.grad = None
        torch_loss(a, b, w, 0.25).backward()

    def our_loss():
        a.grad = None
        losses.image_loss(a, b, 0.25)["total"].backward()

    out["image_loss_fwd_bwd_ms_torch_ops"] = timeit(ref_loss)
    out["image_loss_fwd_bwd_ms_ours"] = timeit(our_loss)

    v = torch.rand(32, 32, 32, device=dev, requires_grad=True)

    def ref_tv():
        v.grad = None
        t = v.diff(dim=0).abs().sum() + v.diff(dim=1).abs().sum() + v.diff(dim=2).abs().sum()
        (t / (3 * 31 * 32 * 32)).backward()

    def our_tv():
        v.grad = None
        losses.tv_3d_loss(v, "mean").backward()

    out["tv32_fwd_bwd_ms_torch_ops"] = timeit(ref_tv)
    out["tv32_fwd_bwd_ms_ours"] = timeit(our_tv)

    P = 100000
    shapes = [(P, 3), (P, 1), (P, 3), (P, 4)]
    for name, cls in (("torch", torch.optim.Adam), ("ours", FusedAdam)):
        ps = [torch.nn.Parameter(torch.randn(*s, device=dev)) for s in shapes]
        for p in ps:
            p.grad = torch.randn_like(p)
        opt = cls([{"params": [p], "lr": 1e-3, "name": str(i)} for i, p in enumerate(ps)], lr=0.0, eps=1e-15)
        out[f"adam_step_100k_ms_{name}"] = timeit(opt.step)

    for n in (100000, 500000):
        pts = torch.rand(n, 3, device=dev) * 2 - 1
        out[f"knn3_{n // 1000}k_ms_ours"] = timeit(lambda: distCUDA2(pts), n=10, warm=2)

    # one training iteration on the headline scene (100k Gaussians, 512^2 cone beam, TV on a 32^3 crop)
    scanner = scene.cone_beam_scanner(512)
    cams = [scene.camera_from_view(vw) for vw in scene.make_views(scanner, 8)]
    cloud = scene.make_cloud(100000, seed=0)
    opt_args = types.SimpleNamespace(
        position_lr_init=2e-4, position_lr_final=2e-5, position_lr_max_steps=30000,
        density_lr_init=1e-2, density_lr_final=1e-3, density_lr_max_steps=30000,
        scaling_lr_init=5e-3, scaling_lr_final=5e-4, scaling_lr_max_steps=30000,
        rotation_lr_init=1e-3, rotation_lr_final=1e-4, rotation_lr_max_steps=30000)
    gm = GaussianModel((0.001, 1.0))
    gm.create_from_pcd(cloud.means, np.maximum(cloud.density, 1e-3), 1.0)
    gm.training_setup(opt_args)
    pipe = types.SimpleNamespace(compute_cov3D_python=False, debug=False)
    with torch.no_grad():
        gts = [render(c, gm, pipe)["render"] * 0.9 for c in cams]
    it = [0]

    def train_iter():
        i = it[0] = it[0] + 1
        gm.update_learning_rate(i)
        pkg = render(cams[i % 8], gm, pipe)
        loss = losses.image_loss(pkg["render"], gts[i % 8], 0.25)["total"]
        vol = query(gm, [0.1, 0.0, -0.1], [32, 32, 32], [0.25, 0.25, 0.25], pipe)["vol"]
        loss = loss + 0.05 * losses.tv_3d_loss(vol, "mean")
        loss.backward()
        with torch.no_grad():
            vis = pkg["visibility_filter"]
            gm.update_max_radii(pkg["radii"], vis)
            gm.add_densification_stats(pkg["viewspace_points"], vis)
        gm.optimizer.step()
        gm.optimizer.zero_grad(set_to_none=True)

    out["train_iteration_100k_512_ms_ours"] = timeit(train_iter, n=40, warm=5)

    def densify():
        with torch.no_grad():
            gm.xyz_gradient_accum = torch.rand_like(gm.xyz_gradient_accum) * 1e-3
            gm.denom = torch.ones_like(gm.denom)
            gm.densify_and_prune(9e-4, 1e-5, None, 0.9, 500000, 0.01,
                                 torch.tensor([[-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]], device=dev))

    t = timeit(densify, n=5, warm=1)
    out["densify_and_prune_ms_ours"] = t
    out["gaussians_after_densify"] = int(gm.get_xyz.shape[0])
    print(json.dumps(out))


if __name__ == "__main__":
    main()
