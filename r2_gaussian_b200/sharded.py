"""Gaussian-sharded multi-GPU projection / voxelization.

X-ray accumulation is a plain sum over Gaussians, so the cloud is partitioned by index across ranks
(one process per GPU); every rank renders its partial detector image (or volume) with the single-GPU
kernels and one all-reduce(sum) makes the full image available on every rank -- which is also all the
backward pass needs: the loss and dL/dimage are then replicated by construction and each rank
back-propagates into its own shard with no further communication.

The exchange step is `torch.distributed.all_reduce` (NCCL over NVLink on GPUs; gloo in the CPU tests) or,
after `enable_peer_exchange()`, the one-kernel sum over NVLink peer memory of `r2_gaussian_b200.peer`
(rank-ordered, bitwise identical on all ranks).  The per-rank renderer is injectable so the host logic
(partition, reduction, bookkeeping) is testable without a GPU.
"""
from __future__ import annotations

from typing import Callable

import torch
import torch.distributed as dist


def shard_bounds(P: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous index partition: rank r owns [P*r//world, P*(r+1)//world)."""
    return (P * rank) // world, (P * (rank + 1)) // world


def world_info() -> tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


_PEER = {"on": False, "reducers": {}}
_SHARD = {"on": False, "group": None}


def enable(group=None, on: bool = True):
    """Declare that THIS process group runs Gaussian-sharded: every rank holds a slice of one cloud and render() /
    query() must sum their images / volumes over the ranks.  Sharding is an explicit opt-in (the trainer and bench.py
    call this): a data-parallel or multi-scene job that merely has torch.distributed initialised keeps per-rank images."""
    _SHARD["on"], _SHARD["group"] = bool(on), (group if on else None)


def enabled() -> bool:
    return bool(_SHARD["on"]) and dist.is_available() and dist.is_initialized()


def sharded_sum(x: torch.Tensor) -> torch.Tensor:
    """all_reduce_sum over the sharding group when sharding is enabled, identity otherwise (render() / query())."""
    return all_reduce_sum(x, _SHARD["group"]) if enabled() else x


def sharded_sum_(x: torch.Tensor) -> torch.Tensor:
    """In-place, autograd-free form of sharded_sum (train_step.NativeTrainStep): x becomes the sum over the ranks."""
    if not enabled() or dist.get_world_size(_SHARD["group"]) == 1:
        return x
    group = _SHARD["group"]
    if _PEER["on"] and x.is_cuda and x.dtype == torch.float32 and x.is_contiguous():
        red = _peer_reducer(x, group)
        red.partial().view_as(x).copy_(x)
        return red.reduce(x)
    dist.all_reduce(x, op=dist.ReduceOp.SUM, group=group)
    return x


def check_peer_exchange():
    """Raise if any peer-memory reduction since the last check gave up waiting for a peer (the kernel sets a status word
    instead of hanging the GPU, and then sums whatever the late peer's buffer held).  Synchronises; the trainer calls it
    where it synchronises anyway (loss read-outs, saves, evaluations)."""
    bad = [k for k, r in _PEER["reducers"].items() if not r.ok()]
    if bad:
        raise RuntimeError(f"r2x_peer_allreduce_sum: a peer did not arrive within the time-out for buffers {bad}; the "
                           "summed images since the last check are invalid (rank skew, e.g. rank-0-only I/O without a barrier)")


def enable_peer_exchange(on: bool = True):
    """Route all_reduce_sum() of CUDA float32 tensors through PeerReducer (single node, one process per GPU).
    Reducers are created on first use per tensor size -- collectively, so every rank must reduce the same sizes in
    the same order (true for render()/query(), whose outputs are replicated shapes)."""
    _PEER["on"] = bool(on)
    if not on:
        for r in _PEER["reducers"].values():
            r.close()
        _PEER["reducers"].clear()


def _peer_reducer(x, group):
    key = (x.numel(), x.device.index, id(group))
    red = _PEER["reducers"].get(key)
    if red is None:
        from .peer import PeerReducer
        red = _PEER["reducers"][key] = PeerReducer(x.numel(), x.device, group)
    return red


class _AllReduceSum(torch.autograd.Function):
    """y = sum over ranks of x.  Backward: every rank already holds the full dL/dy, and dy/dx_r = I."""

    @staticmethod
    def forward(ctx, x, group):
        if _PEER["on"] and x.is_cuda and x.dtype == torch.float32:
            red = _peer_reducer(x, group)
            red.partial().view_as(x).copy_(x)
            return red.reduce(torch.empty_like(x, memory_format=torch.contiguous_format))
        y = x.contiguous().clone()
        dist.all_reduce(y, op=dist.ReduceOp.SUM, group=group)
        return y

    @staticmethod
    def backward(ctx, g):
        return g, None


def all_reduce_sum(x: torch.Tensor, group=None) -> torch.Tensor:
    """Differentiable sum over ranks (identity when not distributed)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return x
    return _AllReduceSum.apply(x, group)


class ShardedProjector:
    """Forward-only sharded projector: `partial_fn(*args) -> tensor` renders this rank's shard into a
    tensor; `__call__` returns the all-reduced result (in place on the partial tensor)."""

    def __init__(self, partial_fn: Callable[..., torch.Tensor], group=None):
        self.partial_fn = partial_fn
        self.group = group

    def __call__(self, *args, **kw) -> torch.Tensor:
        part = self.partial_fn(*args, **kw)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(part, op=dist.ReduceOp.SUM, group=self.group)
        return part


def sharded_render(render_fn: Callable[..., dict], *args, group=None, **kw) -> dict:
    """Wrap a single-GPU `render()`-style function (returning the reference's dict with key "render")
    so that the image is the sum over ranks; the other entries stay per-shard
    (viewspace_points / visibility_filter / radii describe this rank's Gaussians only)."""
    out = render_fn(*args, **kw)
    out = dict(out)
    out["render"] = all_reduce_sum(out["render"], group)
    return out


def sharded_query(query_fn: Callable[..., dict], *args, group=None, **kw) -> dict:
    out = dict(query_fn(*args, **kw))
    out["vol"] = all_reduce_sum(out["vol"], group)
    return out


# ---- Gaussian-sharded training: host-side helpers -----------------------------------------------------------
def shard_init_points(points, dist2, rank: int, world: int):
    """Rank r's contiguous slice of an initial cloud [N,4] (x,y,z,density) and of the mean squared 3-NN distances
    computed on the FULL cloud (so the initial scales are independent of the number of ranks)."""
    lo, hi = shard_bounds(len(points), rank, world)
    return points[lo:hi], (None if dist2 is None else dist2[lo:hi])


def merge_point_clouds(parts: list) -> dict:
    """Concatenate per-rank `save_ply` dictionaries (xyz / density / scale / rotation / scale_bound) in rank order
    into the single point cloud the reference's `test.py` expects."""
    import numpy as np
    if not parts:
        raise ValueError("merge_point_clouds: nothing to merge")
    out = {k: np.concatenate([np.asarray(p[k]) for p in parts], axis=0) for k in ("xyz", "density", "scale", "rotation")}
    out["scale_bound"] = parts[0]["scale_bound"]
    return out


def gather_point_cloud(gaussians, group=None):
    """All ranks call; rank 0 receives the merged dictionary (others None).  Uses object collectives: meant for the
    infrequent save / export steps, not for the training loop."""
    part = {"xyz": gaussians._xyz.detach().cpu().numpy(), "density": gaussians._density.detach().cpu().numpy(),
            "scale": gaussians._scaling.detach().cpu().numpy(), "rotation": gaussians._rotation.detach().cpu().numpy(),
            "scale_bound": gaussians.scale_bound}
    rank, world = world_info()
    if world == 1:
        return part
    parts = [None] * world
    dist.all_gather_object(parts, part, group=group)
    return merge_point_clouds(parts) if rank == 0 else None
