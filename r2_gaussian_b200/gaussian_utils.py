"""Small numeric helpers used by GaussianModel (same names and results as
`r2_gaussian/utils/gaussian_utils.py:5-90`)."""
from __future__ import annotations

import math

import torch


def inverse_softplus(x, beta: float = 1.0):
    return torch.log(torch.exp(beta * x) - 1) / beta


def inverse_sigmoid(x):
    return torch.log(x / (1 - x))


def get_expon_lr_func(lr_init, lr_final, lr_delay_steps=0, lr_delay_mult=1.0, max_steps=1000000):
    """Log-linear interpolation lr_init -> lr_final over max_steps, optionally eased in over lr_delay_steps."""
    log_a, log_b = (math.log(lr_init), math.log(lr_final)) if lr_init > 0 and lr_final > 0 else (None, None)

    def schedule(step):
        if step < 0 or (lr_init == 0.0 and lr_final == 0.0):
            return 0.0
        ease = 1.0
        if lr_delay_steps > 0:
            ease = lr_delay_mult + (1 - lr_delay_mult) * math.sin(0.5 * math.pi * min(max(step / lr_delay_steps, 0.0), 1.0))
        t = min(max(step / max_steps, 0.0), 1.0)
        return ease * math.exp(log_a * (1 - t) + log_b * t)

    return schedule


def build_rotation(q):
    """[N,4] quaternions (r,x,y,z), normalised here -> [N,3,3] rotation matrices."""
    # the norm is spelt out term by term (not q.norm()): the split children's positions depend on these bits, and the
    # reference sums r^2 + x^2 + y^2 + z^2 left to right (utils/gaussian_utils.py:50-54)
    norm = torch.sqrt(q[:, 0] * q[:, 0] + q[:, 1] * q[:, 1] + q[:, 2] * q[:, 2] + q[:, 3] * q[:, 3])
    q = q / norm[:, None]
    r, x, y, z = q.unbind(1)
    rows = (1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
            2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
            2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y))
    return torch.stack(rows, dim=1).reshape(-1, 3, 3)


def build_scaling_rotation(s, q):
    """L = R(q) diag(s): [N,3,3]."""
    return build_rotation(q) * s.unsqueeze(1)


def strip_symmetric(m):
    """Upper triangle (00, 01, 02, 11, 12, 22) of [N,3,3] symmetric matrices -> [N,6]."""
    return torch.stack((m[:, 0, 0], m[:, 0, 1], m[:, 0, 2], m[:, 1, 1], m[:, 1, 2], m[:, 2, 2]), dim=1)
