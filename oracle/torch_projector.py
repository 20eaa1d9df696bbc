"""Pure-PyTorch CPU additive X-ray projector -- the CPU baseline SURVEY.md §8(d) describes.

TEST / BASELINE INFRASTRUCTURE ONLY (like everything under oracle/): imported by tests/ and by bench.py's
cpu_baseline leg, never by the product.  It restates the reference's forward path with vectorised torch ops on the
host cores: per-Gaussian preprocess (RAS/forward.cu:77-289: near cull, Sigma_3 = R S^2 R^T from the un-normalised
quaternion, ray-space covariance through J W, mu in float64, conic, 3-sigma radius, tile rectangle), then a
footprint-restricted evaluation: the (Gaussian, tile) instances are expanded, every instance evaluates the 256
pixels of its tile (RAS/forward.cu:294-395: skip power > 0, skip alpha < 1e-5) and the contributions are
`index_add_`-ed into the image.  Not bit-exact (torch orders float operations differently); checked against the C
oracle to 1e-4 of the image scale in tests/test_oracle_cpu.py.
"""
from __future__ import annotations

import math

import torch

TILE = 16


def _preprocess(means, dens, scales, rots, view_t, proj_t, W, H, tanfovx, tanfovy, mode, scale_modifier=1.0):
    f32, f64 = torch.float32, torch.float64
    V = view_t.to(f32)                     # transposed matrices: row vector convention, p_view = [p 1] @ V
    PV = proj_t.to(f32)
    t = means @ V[:3, :3] + V[3, :3]
    hom = means @ PV[:3, :] + PV[3, :]
    keep = t[:, 2] > 0.2
    pw = 1.0 / (hom[:, 3] + 1e-7)
    ndc_x, ndc_y = hom[:, 0] * pw, hom[:, 1] * pw
    # Sigma_3
    r, x, y, z = rots.unbind(1)
    R = torch.stack((1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                     2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                     2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)), dim=1).reshape(-1, 3, 3)
    M = R * (scale_modifier * scales).unsqueeze(1)
    sigma = M @ M.transpose(1, 2)
    # ray-space covariance
    fx, fy = W / (2.0 * tanfovx), H / (2.0 * tanfovy)
    tx, ty, tz = t.unbind(1)
    zero, one = torch.zeros_like(tz), torch.ones_like(tz)
    if mode == 0:
        J = torch.stack((fx * one, zero, zero, zero, fy * one, zero, zero, zero, one), dim=1).reshape(-1, 3, 3)
    else:
        limx, limy = 1.3 * tanfovx, 1.3 * tanfovy
        tzs = torch.where(keep, tz, one)
        cx = tzs * torch.clamp(tx / tzs, -limx, limx)
        cy = tzs * torch.clamp(ty / tzs, -limy, limy)
        l = torch.sqrt(cx * cx + cy * cy + tzs * tzs)
        J = torch.stack((fx / tzs, zero, -fx * cx / (tzs * tzs), zero, fy / tzs, -fy * cy / (tzs * tzs),
                         cx / l, cy / l, tzs / l), dim=1).reshape(-1, 3, 3)
    A = J @ V[:3, :3].t()
    cov = A @ sigma @ A.transpose(1, 2)
    a, b, c = cov[:, 0, 0], cov[:, 0, 1], cov[:, 0, 2]
    d, e, f = cov[:, 1, 1], cov[:, 1, 2], cov[:, 2, 2]
    det = a * d - b * b
    det3 = a * d * f + 2 * b * c * e - a * e * e - b * b * f - c * c * d
    keep = keep & (det != 0)
    det_s = torch.where(det != 0, det, one)
    musq = det3.to(f64) * (2.0 * math.pi) / det_s.to(f64)
    mu = torch.where(musq > 0, torch.sqrt(musq.clamp_min(0)), torch.zeros_like(musq)).to(f32)
    conic = torch.stack((d / det_s, -b / det_s, a / det_s), dim=1)
    mid = 0.5 * (a + d)
    lam = mid + torch.sqrt(torch.clamp(mid * mid - det, min=0.1))
    radius = torch.ceil(3.0 * torch.sqrt(lam.clamp_min(0)))
    radius = torch.where(keep, radius, torch.zeros_like(radius))      # masked rows may hold NaN
    pix_x = (((ndc_x.to(f64) + 1.0) * W - 1.0) * 0.5).to(f32)
    pix_y = (((ndc_y.to(f64) + 1.0) * H - 1.0) * 0.5).to(f32)
    gx, gy = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE
    tr = lambda v, hi: torch.clamp(torch.nan_to_num(v, nan=0.0, posinf=1e9, neginf=-1e9).to(torch.int64), 0, hi)
    x0, y0 = tr((pix_x - radius) / TILE, gx), tr((pix_y - radius) / TILE, gy)
    x1, y1 = tr((pix_x + radius + TILE - 1) / TILE, gx), tr((pix_y + radius + TILE - 1) / TILE, gy)
    ntiles = torch.where(keep, (x1 - x0) * (y1 - y0), torch.zeros_like(x0))
    return {"xy": torch.stack((pix_x, pix_y), 1), "conic": conic, "w": dens.reshape(-1) * mu, "ntiles": ntiles,
            "x0": x0, "y0": y0, "wt": x1 - x0, "radius": radius.to(torch.int32)}


def project(means, dens, scales, rots, view_t, proj_t, W, H, tanfovx, tanfovy, mode, scale_modifier=1.0,
            chunk_instances=8192):
    """[1,H,W] float32 projection on the CPU.  Inputs are torch (or array-like) float32; matrices transposed as the
    extension expects them."""
    as_t = lambda a: torch.as_tensor(a, dtype=torch.float32)
    means, dens, scales, rots = as_t(means).reshape(-1, 3), as_t(dens), as_t(scales).reshape(-1, 3), as_t(rots).reshape(-1, 4)
    g = _preprocess(means, dens, scales, rots, as_t(view_t).reshape(4, 4), as_t(proj_t).reshape(4, 4), W, H,
                    float(tanfovx), float(tanfovy), int(mode), float(scale_modifier))
    image = torch.zeros(H * W, dtype=torch.float32)
    ids = torch.nonzero(g["ntiles"] > 0).squeeze(1)
    if ids.numel() == 0:
        return image.view(1, H, W)
    nt = g["ntiles"][ids]
    owner = torch.repeat_interleave(ids, nt)                                  # instance -> Gaussian
    start = torch.cumsum(nt, 0) - nt
    k = torch.arange(owner.numel()) - torch.repeat_interleave(start, nt)       # index of the tile inside the rectangle
    wt = g["wt"][owner]
    tile_x = g["x0"][owner] + k % wt
    tile_y = g["y0"][owner] + k // wt
    lx = torch.arange(TILE).repeat(TILE)                                       # pixel offsets inside a tile, row-major
    ly = torch.arange(TILE).repeat_interleave(TILE)
    for s in range(0, owner.numel(), chunk_instances):
        o = owner[s:s + chunk_instances]
        px = (tile_x[s:s + chunk_instances] * TILE).unsqueeze(1) + lx           # [n, 256]
        py = (tile_y[s:s + chunk_instances] * TILE).unsqueeze(1) + ly
        dx = g["xy"][o, 0].unsqueeze(1) - px.to(torch.float32)
        dy = g["xy"][o, 1].unsqueeze(1) - py.to(torch.float32)
        con = g["conic"][o]
        power = -0.5 * (con[:, 0:1] * dx * dx + con[:, 2:3] * dy * dy) - con[:, 1:2] * dx * dy
        alpha = g["w"][o].unsqueeze(1) * torch.exp(power)
        ok = (power <= 0) & (alpha >= 1e-5) & (px < W) & (py < H)
        alpha = torch.where(ok, alpha, torch.zeros_like(alpha))
        image.index_add_(0, (py.clamp_max(H - 1) * W + px.clamp_max(W - 1)).reshape(-1), alpha.reshape(-1))
    return image.view(1, H, W)
