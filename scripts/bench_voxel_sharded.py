"""Voxelizer sweep (BASELINE.json configs[4]): 256^3 volume query over 500k Gaussians, Gaussian-sharded over the
ranks of a torchrun launch, one NCCL all-reduce of the volume per query.  Also runs standalone (1 GPU), where it
times the compiled reference voxelizer (oracle/_ref) next to ours.  Rank 0 prints one JSON line.

    python scripts/bench_voxel_sharded.py
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/bench_voxel_sharded.py
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from r2_gaussian_b200 import scene  # noqa: E402
from r2_gaussian_b200.engine import VoxelEngine  # noqa: E402
from r2_gaussian_b200.sharded import shard_bounds  # noqa: E402


def main():
    rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    P, n = 500000, 256
    cloud = scene.make_cloud(P, kind="trained", seed=1)
    lo, hi = shard_bounds(P, rank, world)
    t = lambda a: torch.tensor(np.ascontiguousarray(a[lo:hi]), device=dev)
    means, dens, scales, rots = t(cloud.means), t(cloud.density), t(cloud.scales), t(cloud.rotations)
    eng = VoxelEngine(hi - lo, (n, n, n), device=dev)
    args = (means, dens, scales, rots, (2.0, 2.0, 2.0), (0.0, 0.0, 0.0))
    eng.fit(*args)
    steps, warm = 20, 3
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step():
        vol = eng.forward(*args)
        if world > 1:
            dist.all_reduce(vol)
        return vol

    for _ in range(warm):
        step()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    for a, b in ev:
        flush.fill_(1)
        a.record()
        step()
        b.record()
    torch.cuda.synchronize()
    ms = torch.tensor([sum(a.elapsed_time(b) for a, b in ev) / steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    out = {"metric": "volume queries/sec (500k Gaussians, 256^3 grid)", "n_gpus": world, "ms_per_query": ms.item(),
           "value": 1e3 / ms.item(), "R_rank0": eng.num_rendered(), "sharding": "gaussians by index + all_reduce(256^3 f32)"}
    if world == 1:
        try:
            import util
            ref = util.ref_lib()
            import ctypes as C
            vol = torch.zeros((n, n, n), dtype=torch.float32, device=dev)
            rx = torch.zeros(P, dtype=torch.int32, device=dev)
            ry, rz = torch.zeros_like(rx), torch.zeros_like(rx)
            fp = lambda x: C.c_void_p(x.data_ptr())
            ref.ref_voxel_forward.restype = C.c_int

            def ref_step():
                vol.zero_()
                return ref.ref_voxel_forward(P, n, n, n, C.c_float(2.0), C.c_float(2.0), C.c_float(2.0), C.c_float(0.0),
                                             C.c_float(0.0), C.c_float(0.0), fp(means), fp(dens), fp(scales),
                                             C.c_float(1.0), fp(rots), None, fp(vol), fp(rx), fp(ry), fp(rz))
            for _ in range(2):
                ref_step()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                ref_step()
            b.record()
            torch.cuda.synchronize()
            out["reference_ms_per_query"] = a.elapsed_time(b) / 5
        except Exception as e:  # the reference build is optional here
            out["reference_ms_per_query"] = f"unavailable: {e}"
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
