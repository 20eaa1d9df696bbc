#!/usr/bin/env python
"""Where one training iteration's time goes (100k Gaussians, 512^2, the secondary block's iteration): wall per
iteration, host enqueue time per iteration (no synchronisation), summed GPU kernel time per iteration and the top
kernels / host ops from torch.profiler.  Prints one JSON object; run on a GPU box."""
import contextlib
import io
import json
import os
import sys
import time
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from r2_gaussian_b200 import losses, scene  # noqa: E402
from r2_gaussian_b200.gaussian_model import GaussianModel  # noqa: E402
from r2_gaussian_b200.render_query import query, render  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    cloud = scene.make_cloud(100000, kind="init", seed=0)
    scanner = scene.cone_beam_scanner(512)
    cams = [scene.camera_from_view(vw, device=dev) for vw in scene.make_views(scanner, 8)]
    opt_args = types.SimpleNamespace(
        position_lr_init=2e-4, position_lr_final=2e-5, position_lr_max_steps=30000,
        density_lr_init=1e-2, density_lr_final=1e-3, density_lr_max_steps=30000,
        scaling_lr_init=5e-3, scaling_lr_final=5e-4, scaling_lr_max_steps=30000,
        rotation_lr_init=1e-3, rotation_lr_final=1e-4, rotation_lr_max_steps=30000)
    gm = GaussianModel((0.001, 1.0))
    with contextlib.redirect_stdout(io.StringIO()):
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
        vol_ = query(gm, [0.1, 0.0, -0.1], [32, 32, 32], [0.25, 0.25, 0.25], pipe)["vol"]
        loss = loss + 0.05 * losses.tv_3d_loss(vol_, "mean")
        loss.backward()
        with torch.no_grad():
            vis = pkg["visibility_filter"]
            gm.update_max_radii(pkg["radii"], vis)
            gm.add_densification_stats(pkg["viewspace_points"], vis)
        gm.optimizer.step()
        gm.optimizer.zero_grad(set_to_none=True)

    for _ in range(10):
        train_iter()
    torch.cuda.synchronize()
    n = 50
    t0 = time.perf_counter()
    for _ in range(n):
        train_iter()
    t_enq = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t0
    out = {"wall_ms_per_iteration": t_wall / n * 1e3, "host_enqueue_ms_per_iteration": t_enq / n * 1e3,
           "fused_activations": os.environ.get("R2X_FUSED_ACTIVATIONS", "default")}
    # the same iteration as a fixed launch sequence (what trainer.py runs)
    from r2_gaussian_b200.train_step import NativeTrainStep
    native = NativeTrainStep(gm, 0.25, 0.05, [32, 32, 32], [0.25, 0.25, 0.25])

    def native_iter():
        i = it[0] = it[0] + 1
        gm.update_learning_rate(i)
        native(cams[i % 8], gts[i % 8], (0.1, 0.0, -0.1))

    for _ in range(10):
        native_iter()
    native.flush()
    torch.cuda.synchronize()
    nn = 300
    t0 = time.perf_counter()
    for _ in range(nn):
        native_iter()
    t_enq = time.perf_counter() - t0
    native.flush()
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t0
    out["native_wall_ms_per_iteration"] = t_wall / nn * 1e3
    out["native_host_enqueue_ms_per_iteration"] = t_enq / nn * 1e3
    out["native_repeats"] = native.repeats
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(nn):
        native_iter()
    ev1.record()
    native.flush()
    torch.cuda.synchronize()
    out["native_gpu_ms_per_iteration_events"] = ev0.elapsed_time(ev1) / nn
    if os.environ.get("TRAIN_CPROFILE"):      # where the HOST time of an iteration goes, Python frames included
        import cProfile
        import pstats
        pr = cProfile.Profile()
        pr.enable()
        for _ in range(n):
            train_iter()
        pr.disable()
        torch.cuda.synchronize()
        st = pstats.Stats(pr)
        rows = sorted(st.stats.items(), key=lambda kv: kv[1][3], reverse=True)[:40]
        out["cprofile_top_cumulative"] = [
            {"fn": f"{os.path.basename(k[0])}:{k[1]}:{k[2]}", "calls_per_iteration": v[1] / n,
             "self_us_per_iteration": v[2] / n * 1e6, "cum_us_per_iteration": v[3] / n * 1e6} for k, v in rows]
    from torch.profiler import ProfilerActivity, profile
    m = 20
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(m):
            train_iter()
        torch.cuda.synchronize()
    ka = prof.key_averages()
    def cuda_t(e):
        return getattr(e, "self_device_time_total", getattr(e, "self_cuda_time_total", 0.0))
    kern = sorted(ka, key=cuda_t, reverse=True)
    out["gpu_kernel_ms_per_iteration"] = sum(cuda_t(e) for e in ka) / m / 1e3
    out["gpu_launches_per_iteration"] = sum(e.count for e in ka if cuda_t(e) > 0) / m
    out["top_gpu"] = [{"name": e.key[:70], "us_per_iteration": cuda_t(e) / m, "calls_per_iteration": e.count / m}
                      for e in kern[:25] if cuda_t(e) > 0]
    cpu = sorted(ka, key=lambda e: e.self_cpu_time_total, reverse=True)
    out["top_host"] = [{"name": e.key[:70], "us_per_iteration": e.self_cpu_time_total / m, "calls_per_iteration": e.count / m}
                       for e in cpu[:25]]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
