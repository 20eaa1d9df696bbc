"""The C-ABI library loads without a GPU and exports exactly what include/r2x.h declares; the ctypes
prototypes in r2_gaussian_b200/_lib.py agree with the header (names and argument counts)."""
import ctypes
import os
import re

import util
from r2_gaussian_b200 import _lib

HDR = os.path.join(util.ROOT, "include", "r2x.h")


def _declared():
    src = open(HDR).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decls = {}
    for m in re.finditer(r"\b(?:int|size_t|const char\s*\*)\s+(r2x_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        name, args = m.group(1), m.group(2).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        decls[name] = n
    return decls


def test_header_declares_the_reference_entry_points():
    d = _declared()
    for must in ["r2x_raster_forward", "r2x_raster_backward", "r2x_mark_visible", "r2x_voxel_forward",
                 "r2x_voxel_backward", "r2x_raster_forward_async", "r2x_voxel_forward_async"]:
        assert must in d


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH) if os.path.exists(_lib.LIB_PATH) else _lib.load()
    for name in _declared():
        assert hasattr(lib, name), f"libr2xray.so lacks {name}"


def test_ctypes_prototypes_match_header():
    d = _declared()
    assert set(d) == set(_lib.PROTOTYPES), set(d) ^ set(_lib.PROTOTYPES)
    for name, n in d.items():
        assert len(_lib.PROTOTYPES[name][1]) == n, f"{name}: header has {n} args, ctypes {len(_lib.PROTOTYPES[name][1])}"


def test_size_queries_and_error_path_without_gpu():
    lib = _lib.load()
    assert lib.r2x_version() >= 100
    assert lib.r2x_raster_geom_bytes(1000) >= 1000 * (32 + 16 + 12 + 8)
    assert lib.r2x_binning_bytes(5000) >= 5000 * 28
    assert lib.r2x_raster_image_bytes(1000, 512, 512) >= 1024 * 8
    assert lib.r2x_voxel_image_bytes(1000, 256, 256, 256) >= 32768 * 8
    # invalid arguments are rejected before any CUDA call
    rc = lib.r2x_raster_forward_async(None, 10, 0, 16, None, None, None, 1.0, None, None, None, None, None, 1.0, 1.0, 0, 1,
                                      None, None, None, None, None, 0, None)
    assert rc != 0 and b"bad" in lib.r2x_last_error()
