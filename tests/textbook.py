"""Independent float64 NumPy statement of the R2-Gaussian projection / voxelization MATH (paper
equations as summarised in SURVEY.md Appendix A), written without looking at how the oracle orders its
operations.  Used on the CPU to catch formula errors in oracle/r2_oracle.c: agreement is to ~1e-5, not
bit-exact, and only away from the discontinuities (ceil of the radius, tile rectangles, alpha cut)."""
import numpy as np


def quat_to_rot(q):
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = np.empty((q.shape[0], 3, 3))
    R[:, 0, 0] = 1 - 2 * (y * y + z * z); R[:, 0, 1] = 2 * (x * y - r * z); R[:, 0, 2] = 2 * (x * z + r * y)
    R[:, 1, 0] = 2 * (x * y + r * z); R[:, 1, 1] = 1 - 2 * (x * x + z * z); R[:, 1, 2] = 2 * (y * z - r * x)
    R[:, 2, 0] = 2 * (x * z - r * y); R[:, 2, 1] = 2 * (y * z + r * x); R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def sigma3(scales, rots, mod=1.0):
    R = quat_to_rot(rots.astype(np.float64))
    S2 = (mod * scales.astype(np.float64)) ** 2
    return np.einsum("nij,nj,nkj->nik", R, S2, R)


def project(means, scales, rots, view_t, proj_t, W, H, tanfovx, tanfovy, mode):
    """Returns dict(xy, depth, cov2 (a,b,d), conic (A,B,C), mu, radius_float)."""
    means = means.astype(np.float64)
    V = view_t.astype(np.float64).T  # proper world->view 4x4
    Pm = proj_t.astype(np.float64).T
    ph = np.c_[means, np.ones(len(means))]
    t = (ph @ V.T)[:, :3]
    hom = ph @ Pm.T
    ndc = hom[:, :2] / (hom[:, 3:4] + 1e-7)
    xy = np.stack([((ndc[:, 0] + 1) * W - 1) * 0.5, ((ndc[:, 1] + 1) * H - 1) * 0.5], axis=1)
    fx, fy = W / (2 * tanfovx), H / (2 * tanfovy)
    Rw = V[:3, :3]
    Sig = sigma3(scales, rots)
    n = len(means)
    J = np.zeros((n, 3, 3))
    if mode == 0:
        J[:, 0, 0] = fx; J[:, 1, 1] = fy; J[:, 2, 2] = 1.0
    else:
        tz = t[:, 2]
        tx = np.clip(t[:, 0] / tz, -1.3 * tanfovx, 1.3 * tanfovx) * tz
        ty = np.clip(t[:, 1] / tz, -1.3 * tanfovy, 1.3 * tanfovy) * tz
        l = np.sqrt(tx * tx + ty * ty + tz * tz)
        J[:, 0, 0] = fx / tz; J[:, 0, 2] = -fx * tx / tz ** 2
        J[:, 1, 1] = fy / tz; J[:, 1, 2] = -fy * ty / tz ** 2
        J[:, 2, 0] = tx / l; J[:, 2, 1] = ty / l; J[:, 2, 2] = tz / l
    A = J @ Rw[None]
    cov = A @ Sig @ np.transpose(A, (0, 2, 1))
    a, b, d = cov[:, 0, 0], cov[:, 0, 1], cov[:, 1, 1]
    det2 = a * d - b * b
    det3 = np.linalg.det(cov)
    mu = np.sqrt(np.maximum(2 * np.pi * det3 / det2, 0))
    conic = np.stack([d / det2, -b / det2, a / det2], axis=1)
    mid = 0.5 * (a + d)
    lam = mid + np.sqrt(np.maximum(0.1, mid * mid - det2))
    return dict(xy=xy, depth=t[:, 2], cov2=np.stack([a, b, d], 1), conic=conic, mu=mu, radius=3 * np.sqrt(lam))


def render_bruteforce(xy, conic, w, W, H, mask=None, thr=1e-5):
    """Sum over ALL Gaussians at every pixel of w*exp(power) with the reference's two skip rules.
    (No tile rectangles: equals the tiled result wherever contributions outside a Gaussian's
    rectangle are below the cut anyway.)"""
    ys, xs = np.mgrid[0:H, 0:W]
    img = np.zeros((H, W))
    idx = range(len(xy)) if mask is None else np.nonzero(mask)[0]
    for g in idx:
        dx = xy[g, 0] - xs; dy = xy[g, 1] - ys
        p = -0.5 * (conic[g, 0] * dx * dx + conic[g, 2] * dy * dy) - conic[g, 1] * dx * dy
        al = w[g] * np.exp(np.minimum(p, 0))
        img += np.where((p <= 0) & (al >= thr), al, 0.0)
    return img


def voxel_conic(scales, rots, dvox):
    Sig = sigma3(scales, rots)
    D = np.diag(1.0 / np.asarray(dvox, dtype=np.float64))
    cov = D[None] @ Sig @ D[None]
    return np.linalg.inv(cov)
