"""Device-side row selection and compaction (libr2xray: r2x_mask_select / r2x_gather_rows).

Used by `GaussianModel` for clone / split / prune instead of boolean-mask indexing (`t[mask]` = nonzero + gather with a
host synchronisation per tensor in the reference, gaussian_model.py:335-403): a mask becomes a stable index list plus a
device-side count, and ONE launch gathers every per-Gaussian tensor through it.
"""
from __future__ import annotations

import ctypes as C

import torch

from ._lib import GatherDesc, check, load


def select_rows(mask: torch.Tensor):
    """mask: bool/uint8 [n] on CUDA -> (idx int32[n] whose first `count` entries are the selected rows in ascending
    order, count uint32-as-int32 [1] on the device).  No host synchronisation."""
    if mask.device.type != "cuda":
        raise RuntimeError("select_rows: the mask must be a CUDA tensor (no CPU fallback)")
    lib = load()
    m = mask.reshape(-1)
    m = (m if m.dtype in (torch.bool, torch.uint8) else (m != 0)).contiguous()
    n = int(m.numel())
    dev = m.device
    idx = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    count = torch.empty(1, dtype=torch.int32, device=dev)
    nbytes = lib.r2x_mask_select_scratch_bytes(n)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.r2x_mask_select(torch.cuda.current_stream(dev).cuda_stream, n, m.data_ptr() if n else None,
                                 idx.data_ptr(), count.data_ptr(), scratch.data_ptr(), nbytes)
    check(rc, "r2x_mask_select")
    return idx, count


def read_counts(*counts) -> list[int]:
    """The ONE host round trip: several device-side counts in a single read."""
    return [int(v) for v in torch.cat([c.reshape(1) for c in counts]).tolist()]


def gather_rows(specs, select, nsel: int):
    """specs: list of (src0 [n0, w] or [n0], src1 or None); returns the gathered tensors [nsel, w] (or [nsel]).
    Source row s is src0[s] for s < n0, otherwise src1[s - n0] (zeros when src1 is None); `select` int32[>= nsel] or
    None for the identity.  All tensors float32 CUDA; one kernel launch for all of them."""
    lib = load()
    if not specs:
        return []
    dev = specs[0][0].device
    descs, outs, keep = [], [], []
    for src0, src1 in specs:
        a = src0.detach()
        a = a if a.is_contiguous() else a.contiguous()
        width = 1 if a.dim() == 1 else int(a.numel() // max(a.shape[0], 1)) if a.shape[0] else int(torch.tensor(a.shape[1:]).prod())
        b = None
        if src1 is not None and src1.numel():
            b = src1.detach()
            b = b if b.is_contiguous() else b.contiguous()
            if b.dtype != torch.float32:
                raise RuntimeError("gather_rows: float32 tensors only")
        if a.dtype != torch.float32:
            raise RuntimeError("gather_rows: float32 tensors only")
        out = torch.empty((nsel,) + tuple(a.shape[1:]), dtype=torch.float32, device=dev)
        descs.append((a, b, out, int(a.shape[0]), width))
        outs.append(out)
        keep += [a, b]
    if nsel > 0:
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            for i in range(0, len(descs), 16):
                chunk = descs[i:i + 16]
                arr = (GatherDesc * len(chunk))()
                for k, (a, b, out, n0, width) in enumerate(chunk):
                    arr[k].src0 = a.data_ptr() if n0 else None
                    arr[k].src1 = b.data_ptr() if b is not None else None
                    arr[k].dst = out.data_ptr()
                    arr[k].n0 = n0
                    arr[k].width = width
                rc = lib.r2x_gather_rows(stream, len(chunk), C.cast(arr, C.c_void_p),
                                         select.data_ptr() if select is not None else None, int(nsel))
                check(rc, "r2x_gather_rows")
    return outs
