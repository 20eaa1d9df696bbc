"""Debug aid kept from the reference: with `debug=True` the argument tuple of a failing native call is
written to `snapshot_fw.dump` / `snapshot_bw.dump` (PYX/rasterization.py:80-93, :156-175)."""
from __future__ import annotations

import torch


def _cpu_copy(args):
    return tuple(a.detach().cpu().clone() if isinstance(a, torch.Tensor) else a for a in args)


def call_with_snapshot(fn, args, debug: bool, dump_path: str, what: str):
    if not debug:
        return fn(*args)
    saved = _cpu_copy(args)  # taken before the call so a crash cannot corrupt them
    try:
        return fn(*args)
    except Exception:
        torch.save(saved, dump_path)
        print(f"\nAn error occured in {what}. Writing {dump_path} for debugging.\n")
        raise
