"""Adam with the update of every parameter group applied by one CUDA launch (r2x_adam_step).

`FusedAdam` is a `torch.optim.Adam`: same constructor, `param_groups`, per-parameter `state` (`step`, `exp_avg`,
`exp_avg_sq`), `state_dict()` / `load_state_dict()` -- so the optimizer surgery the reference performs when it
densifies and prunes (`gaussian_model.py:320-403`) and its checkpoints (`:79-110`) work unchanged -- only
`step()` is replaced.  Supported configuration = the reference's (`gaussian_model.py:216`): amsgrad off, no
weight decay, maximize off, float32 CUDA parameters.
"""
from __future__ import annotations

import ctypes as C

import torch

from ._lib import AdamGroup, check, load


class FusedAdam(torch.optim.Adam):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, **kw):
        for k in ("weight_decay", "amsgrad", "maximize"):
            if kw.get(k):
                raise RuntimeError(f"FusedAdam: {k} is not supported")
        kw.pop("fused", None), kw.pop("foreach", None)
        super().__init__(params, lr=lr, betas=betas, eps=eps, **kw)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = load()
        buckets = {}
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("FusedAdam: parameters must be contiguous float32 CUDA tensors")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                key = (p.device, int(st["step"].item()), float(b1), float(b2), float(group["eps"]))
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                buckets.setdefault(key, []).append((p, g, st, float(group["lr"])))
        for (dev, step, b1, b2, eps), items in buckets.items():
            with torch.cuda.device(dev):
                stream = torch.cuda.current_stream(dev).cuda_stream
                for i in range(0, len(items), 8):
                    chunk = items[i:i + 8]
                    arr = (AdamGroup * len(chunk))()
                    for k, (p, g, st, lr) in enumerate(chunk):
                        arr[k].param = p.data_ptr()
                        arr[k].grad = g.data_ptr()
                        arr[k].exp_avg = st["exp_avg"].data_ptr()
                        arr[k].exp_avg_sq = st["exp_avg_sq"].data_ptr()
                        arr[k].numel = p.numel()
                        arr[k].lr = lr
                    rc = lib.r2x_adam_step(stream, len(chunk), C.cast(arr, C.c_void_p), b1, b2, eps, step)
                    check(rc, "r2x_adam_step")
        return loss
