"""World-size-2 gloo test of the Gaussian-sharded path on CPU: partition + all-reduce(sum) of the
partial detector images / volumes reproduces the unsharded result, and the differentiable all-reduce
passes gradients straight through.  The per-rank renderer is the CPU oracle (test infrastructure) -- the
host logic under test is r2_gaussian_b200.sharded."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import util
from r2_gaussian_b200 import scene
from r2_gaussian_b200.sharded import ShardedProjector, all_reduce_sum, shard_bounds


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cloud, view = util.case("cone_trained_small")
        lo, hi = shard_bounds(cloud.P, rank, world)
        shard = scene.Cloud(cloud.means[lo:hi], cloud.scales[lo:hi], cloud.rotations[lo:hi], cloud.density[lo:hi])

        def partial_image():
            return torch.from_numpy(util.oracle_raster_forward(shard, view)["image"].copy())

        img = ShardedProjector(partial_image)()
        nV, sV, c = (16, 16, 16), (2.0, 2.0, 2.0), (0.0, 0.0, 0.0)
        vol = ShardedProjector(lambda: torch.from_numpy(util.oracle_voxel_forward(shard, nV, sV, c)["vol"].copy()))()
        # differentiable all-reduce: y = sum_r x_r ; dL/dx_r = dL/dy on every rank
        x = torch.full((4,), float(rank + 1), requires_grad=True)
        y = all_reduce_sum(x * 2.0)
        (y * torch.arange(4.0)).sum().backward()
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), img=img.numpy(), vol=vol.numpy(), y=y.detach().numpy(),
                 gx=x.grad.numpy())
    finally:
        dist.destroy_process_group()


def test_sharded_sum_matches_full_render(tmp_path):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    cloud, view = util.case("cone_trained_small")
    full = util.oracle_raster_forward(cloud, view)["image"].astype(np.float64)
    fullv = util.oracle_voxel_forward(cloud, (16, 16, 16), (2.0, 2.0, 2.0), (0.0, 0.0, 0.0))["vol"].astype(np.float64)
    r = [np.load(os.path.join(tmp_path, f"rank{k}.npz")) for k in range(world)]
    np.testing.assert_array_equal(r[0]["img"], r[1]["img"])          # every rank holds the full image
    assert np.abs(r[0]["img"] - full).max() <= 1e-5 * np.abs(full).max()
    assert np.abs(r[0]["vol"] - fullv).max() <= 1e-5 * np.abs(fullv).max()
    np.testing.assert_allclose(r[0]["y"], np.full(4, 2.0 * (1 + 2)))
    for k in range(world):
        np.testing.assert_allclose(r[k]["gx"], 2.0 * np.arange(4.0))
