"""Loss functions of the reference's training step over the fused CUDA kernels (r2x_image_loss, r2x_tv3d_loss).

Same names and meaning as `r2_gaussian/utils/loss_utils.py` (`l1_loss` :37-38, `ssim` :63-72, `tv_3d_loss` :19-34),
plus `image_loss`, which evaluates the combination train.py:118-127 builds (L1 + lambda * (1 - SSIM)) and its
gradient in two kernel launches instead of ~25 torch ops.  CUDA tensors only (no CPU fallback).
"""
from __future__ import annotations

import torch

from ._lib import check, load


def _require_cuda(name, *tensors):
    for t in tensors:
        if not isinstance(t, torch.Tensor) or not t.is_cuda:
            raise RuntimeError(f"{name}: expected CUDA tensors (this build has no CPU fallback)")


def _as_chw(t):
    if t.dim() == 2:
        return t.unsqueeze(0)
    if t.dim() == 4 and t.shape[0] == 1:
        return t[0]
    if t.dim() != 3:
        raise RuntimeError(f"expected an image of shape [C,H,W], [H,W] or [1,C,H,W], got {tuple(t.shape)}")
    return t


class _ImageLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, target, w_l1, w_dssim):
        lib = load()
        img = _as_chw(image).contiguous().float()
        tgt = _as_chw(target).contiguous().float()
        if img.shape != tgt.shape:
            raise RuntimeError(f"image_loss: shapes differ: {tuple(img.shape)} vs {tuple(tgt.shape)}")
        Cn, H, W = (int(v) for v in img.shape)
        need_grad = image.requires_grad
        dev = img.device
        with torch.cuda.device(dev):
            nbytes = int(lib.r2x_image_loss_scratch_bytes(H, W))
            scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            out = torch.empty((Cn, 3), dtype=torch.float32, device=dev)
            grad = torch.empty_like(img) if need_grad else None
            st = torch.cuda.current_stream(dev).cuda_stream
            for c in range(Cn):
                rc = lib.r2x_image_loss(st, H, W, img[c].data_ptr(), tgt[c].data_ptr(), float(w_l1), float(w_dssim),
                                        out[c].data_ptr(), grad[c].data_ptr() if need_grad else None,
                                        scratch.data_ptr(), nbytes)
                check(rc, "r2x_image_loss")
        res = out[0] if Cn == 1 else out.mean(0)
        if need_grad:
            ctx.save_for_backward(grad if Cn == 1 else grad / Cn)
        ctx.shape = image.shape
        loss, l1, ssim = res[2], res[0], res[1]
        ctx.mark_non_differentiable(l1, ssim)
        return loss, l1, ssim

    @staticmethod
    def backward(ctx, g_loss, _g_l1, _g_ssim):
        (grad,) = ctx.saved_tensors
        return (grad * g_loss).reshape(ctx.shape), None, None, None


def image_loss(image, target, lambda_dssim: float = 0.25, lambda_l1: float = 1.0):
    """{"render": L1, "dssim": 1 - SSIM, "total": lambda_l1 * L1 + lambda_dssim * (1 - SSIM)} (train.py:118-127)."""
    _require_cuda("image_loss", image, target)
    total, l1, ssim_v = _ImageLoss.apply(image, target, float(lambda_l1), float(lambda_dssim))
    return {"total": total, "render": l1, "dssim": 1.0 - ssim_v}


def l1_loss(network_output, gt):
    _require_cuda("l1_loss", network_output, gt)
    return _ImageLoss.apply(network_output, gt, 1.0, 0.0)[0]


def ssim(img1, img2, window_size: int = 11, size_average: bool = True):
    if window_size != 11 or not size_average:
        raise RuntimeError("ssim: the fused kernel implements the reference's call (window 11, size_average=True)")
    _require_cuda("ssim", img1, img2)
    return 1.0 - _ImageLoss.apply(img1, img2, 0.0, 1.0)[0]


class _TV3D(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vol, mean):
        lib = load()
        v = vol.contiguous().float()
        if v.dim() != 3:
            raise RuntimeError(f"tv_3d_loss: expected a [nx,ny,nz] volume, got {tuple(v.shape)}")
        nx, ny, nz = (int(s) for s in v.shape)
        dev = v.device
        with torch.cuda.device(dev):
            nbytes = int(lib.r2x_tv3d_scratch_bytes(nx, ny, nz))
            scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            out = torch.empty(1, dtype=torch.float32, device=dev)
            grad = torch.empty_like(v) if vol.requires_grad else None
            rc = lib.r2x_tv3d_loss(torch.cuda.current_stream(dev).cuda_stream, nx, ny, nz, v.data_ptr(), int(bool(mean)),
                                   out.data_ptr(), grad.data_ptr() if grad is not None else None, scratch.data_ptr(),
                                   nbytes)
        check(rc, "r2x_tv3d_loss")
        if grad is not None:
            ctx.save_for_backward(grad)
        return out[0]

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None


def tv_3d_loss(vol, reduction: str = "sum"):
    _require_cuda("tv_3d_loss", vol)
    return _TV3D.apply(vol, reduction == "mean")
