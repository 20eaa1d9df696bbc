"""`simple_knn._C.distCUDA2` over the C ABI (r2x_knn3_mean_dist2).

The reference imports `distCUDA2` at module load (`r2_gaussian/gaussian/gaussian_model.py:21`) and calls it once,
in `create_from_pcd` (`:144-150`), to size the initial Gaussians: the result is, per point, the mean squared
distance to its three nearest other points.  No CPU fallback: a non-CUDA tensor raises.
"""
from __future__ import annotations

import torch

from ._lib import check, load


def distCUDA2(points: torch.Tensor) -> torch.Tensor:
    if not isinstance(points, torch.Tensor) or not points.is_cuda:
        raise RuntimeError("distCUDA2: expected a CUDA tensor (this build has no CPU fallback)")
    if points.dim() != 2 or points.shape[1] != 3:
        raise RuntimeError(f"distCUDA2: expected points of shape [P, 3], got {tuple(points.shape)}")
    lib = load()
    pts = points.detach().contiguous().float()
    P = int(pts.shape[0])
    out = torch.empty(P, dtype=torch.float32, device=pts.device)
    if P == 0:
        return out
    with torch.cuda.device(pts.device):
        nbytes = int(lib.r2x_knn_scratch_bytes(P))
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=pts.device)
        rc = lib.r2x_knn3_mean_dist2(torch.cuda.current_stream(pts.device).cuda_stream, P, pts.data_ptr(),
                                     out.data_ptr(), scratch.data_ptr(), nbytes)
    check(rc, "r2x_knn3_mean_dist2")
    return out
