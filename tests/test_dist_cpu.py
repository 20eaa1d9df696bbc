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
from r2_gaussian_b200.sharded import (ShardedProjector, all_reduce_sum, gather_point_cloud, merge_point_clouds,
                                      shard_bounds, shard_init_points)


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
        # sharded-training helpers: slices of the initial cloud (with the full-cloud 3-NN distances), merged export
        rng = np.random.default_rng(0)
        pts, d2 = rng.random((101, 4)), rng.random(101)
        my_pts, my_d2 = shard_init_points(pts, d2, rank, world)
        from types import SimpleNamespace
        fake = SimpleNamespace(_xyz=torch.from_numpy(my_pts[:, :3]), _density=torch.from_numpy(my_pts[:, 3:4]),
                               _scaling=torch.from_numpy(np.repeat(my_d2[:, None], 3, 1)),
                               _rotation=torch.zeros(len(my_pts), 4), scale_bound=(0.001, 1.0))
        merged = gather_point_cloud(fake)
        assert (merged is None) == (rank != 0)
        if rank == 0:
            assert np.array_equal(merged["xyz"], pts[:, :3]) and np.array_equal(merged["density"], pts[:, 3:4])
            assert np.array_equal(merged["scale"][:, 0], d2) and merged["rotation"].shape == (101, 4)
            assert merged["scale_bound"] == (0.001, 1.0)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), img=img.numpy(), vol=vol.numpy(), y=y.detach().numpy(),
                 gx=x.grad.numpy(), n_pts=np.array([len(my_pts)]))
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
    assert int(r[0]["n_pts"][0]) + int(r[1]["n_pts"][0]) == 101 and int(r[0]["n_pts"][0]) == 50


def test_merge_point_clouds_orders_by_rank():
    a = {"xyz": np.zeros((2, 3)), "density": np.zeros((2, 1)), "scale": np.ones((2, 3)), "rotation": np.ones((2, 4)),
         "scale_bound": None}
    b = {k: (v + 1 if isinstance(v, np.ndarray) else v) for k, v in a.items()}
    m = merge_point_clouds([a, b])
    assert m["xyz"].shape == (4, 3) and m["xyz"][:2].max() == 0 and m["xyz"][2:].min() == 1 and m["scale_bound"] is None
    with pytest.raises(ValueError):
        merge_point_clouds([])


def _worker_opt_in(rank, world, port, q):
    """render()/query() sum over ranks only after sharded.enable(); an initialised group alone changes nothing."""
    import torch.distributed as dist
    from r2_gaussian_b200 import sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x = torch.full((4,), float(rank + 1))
        a = sharded.sharded_sum(x).clone()
        sharded.enable()
        b = sharded.sharded_sum(x).clone()
        sharded.enable(on=False)
        c = sharded.sharded_sum(x).clone()
        q.put((rank, a.tolist(), b.tolist(), c.tolist()))
    finally:
        dist.destroy_process_group()


def test_sharding_is_an_explicit_opt_in():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker_opt_in, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, a, b, c in res:
        assert a == [rank + 1.0] * 4 and c == [rank + 1.0] * 4     # untouched without the opt-in
        assert b == [3.0] * 4                                      # 1 + 2 once sharding is enabled


def _worker_inplace(rank, world, port, q):
    """sharded_sum_ (what the native training step exchanges): in place, identity without the opt-in, and the overflow
    flag that rides behind the image comes back as "any rank overflowed" on every rank."""
    import torch.distributed as dist
    from r2_gaussian_b200 import sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        H = W = 4
        ext = torch.zeros(H * W + 4)
        image, flag = ext[:H * W].view(1, H, W), ext[H * W:H * W + 1]
        status = torch.tensor([100 + rank, 1 if rank == 1 else 0], dtype=torch.int32)   # only rank 1 overflowed
        image.fill_(float(rank + 1))
        before = sharded.sharded_sum_(ext).clone()                  # no opt-in: untouched
        sharded.enable()
        flag.copy_(status[1:2])
        out = sharded.sharded_sum_(ext)
        status[1:2].copy_(flag)
        same_storage = out.data_ptr() == ext.data_ptr()
        sharded.enable(on=False)
        q.put((rank, before[:H * W].tolist(), image.flatten().tolist(), status.tolist(), same_storage))
    finally:
        dist.destroy_process_group()


def test_inplace_exchange_carries_the_overflow_flag():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker_inplace, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, before, image, status, same_storage in res:
        assert before == [rank + 1.0] * 16
        assert image == [3.0] * 16 and same_storage
        assert status == [100 + rank, 1]            # the count stays local, the flag is global


def test_rank_checkpoint_paths():
    from r2_gaussian_b200.trainer import rank_checkpoint_path
    assert rank_checkpoint_path("ckpt/chkpnt100.pth", 0, 1) == "ckpt/chkpnt100.pth"
    assert rank_checkpoint_path("ckpt/chkpnt100.pth", 3, 8) == "ckpt/chkpnt100_rank3.pth"
    assert rank_checkpoint_path("ckpt/chkpnt100_rank0.pth", 5, 8) == "ckpt/chkpnt100_rank5.pth"
