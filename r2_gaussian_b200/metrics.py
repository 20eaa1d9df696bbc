"""Evaluation metrics with the reference's definitions (`r2_gaussian/utils/image_utils.py:19-183`): `mse`, `rmse`,
`psnr` on [b,c,h,w] batches, `metric_vol` (3-D PSNR; slice-wise SSIM averaged over the three axes) and
`metric_proj` (per-projection, each slice normalised by its own maximum).  Evaluation is not on the hot path: SSIM
here is the plain torch formulation (any device), windows as in `loss_utils.py:45-104`."""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


def _window(channel: int, like: torch.Tensor) -> torch.Tensor:
    g = torch.tensor([math.exp(-((x - 5) ** 2) / float(2 * 1.5 ** 2)) for x in range(11)])
    g = (g / g.sum()).unsqueeze(1)
    w = g.mm(g.t()).float().unsqueeze(0).unsqueeze(0)
    return w.expand(channel, 1, 11, 11).contiguous().to(device=like.device, dtype=like.dtype)


def ssim(img1: torch.Tensor, img2: torch.Tensor) -> torch.Tensor:
    """Mean SSIM of [b,c,h,w] (or [c,h,w]) images, 11x11 Gaussian window, zero padding."""
    c = img1.size(-3)
    w = _window(c, img1)
    mu1, mu2 = F.conv2d(img1, w, padding=5, groups=c), F.conv2d(img2, w, padding=5, groups=c)
    s11 = F.conv2d(img1 * img1, w, padding=5, groups=c) - mu1 * mu1
    s22 = F.conv2d(img2 * img2, w, padding=5, groups=c) - mu2 * mu2
    s12 = F.conv2d(img1 * img2, w, padding=5, groups=c) - mu1 * mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    return (((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 * mu1 + mu2 * mu2 + C1) * (s11 + s22 + C2))).mean()


def mse(img1, img2, mask=None):
    if mask is None:
        return ((img1 - img2) ** 2).reshape(img1.shape[0], -1).mean(1, keepdim=True)
    n_channel = img1.shape[1]
    a, b = img1.flatten(1), img2.flatten(1)
    m = mask.flatten(1).repeat(1, n_channel) != 0
    return torch.stack([((a[i, m[i]] - b[i, m[i]]) ** 2).mean(0, keepdim=True) for i in range(a.shape[0])], dim=0)


def rmse(img1, img2, mask=None):
    return mse(img1, img2, mask) ** 0.5


@torch.no_grad()
def psnr(img1, img2, mask=None, pixel_max=1.0):
    out = 10 * torch.log10(pixel_max ** 2 / mse(img1, img2, mask).float())
    if mask is not None and torch.isinf(out).any():
        out = out[~torch.isinf(out)]
    return out


def _slices(vol, axis):
    for i in range(vol.shape[axis]):
        yield vol.select(axis, i)


def _as_tensor(a):
    return torch.from_numpy(np.array(a, copy=True)) if isinstance(a, np.ndarray) else a


@torch.no_grad()
def metric_vol(img1, img2, metric="psnr", pixel_max=1.0):
    """img1 = ground truth.  -> (value, per-axis list or None)."""
    assert metric in ("psnr", "ssim")
    img1, img2 = _as_tensor(img1), _as_tensor(img2)
    if metric == "psnr":
        if pixel_max is None:
            pixel_max = img1.max()
        return (10 * torch.log10(pixel_max ** 2 / torch.mean((img1 - img2) ** 2).float())).item(), None
    per_axis = []
    for axis in (0, 1, 2):
        vals, count = [], 0
        for s1, s2 in zip(_slices(img1, axis), _slices(img2, axis)):
            if s1.max() > 0:
                vals.append(float(ssim(s1[None, None], s2[None, None])))
                count += 1
            else:
                vals.append(0.0)
        per_axis.append(sum(vals) / count)
    return float(np.mean(per_axis)), per_axis


@torch.no_grad()
def metric_proj(img1, img2, metric="psnr", axis=2, pixel_max=1.0):
    """Stack of projections along `axis`; every non-empty slice is normalised by its own maximum first."""
    assert axis in (0, 1, 2, None) and metric in ("psnr", "ssim")
    img1, img2 = _as_tensor(img1), _as_tensor(img2)
    vals, count = [], 0
    for s1, s2 in zip(_slices(img1, axis), _slices(img2, axis)):
        if s1.max() > 0:
            a, b = (s1 / s1.max())[None, None], (s2 / s2.max())[None, None]
            vals.append(float(psnr(a, b, pixel_max=pixel_max)) if metric == "psnr" else float(ssim(a, b)))
            count += 1
        else:
            vals.append(0.0)
    return sum(vals) / count, vals
