"""The two SimpleITK calls the reference's test.py makes (test.py:141-148): GetImageFromArray + WriteImage of a
float volume to .nii.gz.  This stand-in writes a minimal valid NIfTI-1 file (348-byte header + float32 data, gzip)."""
import gzip
import struct

import numpy as np


class _Image:
    def __init__(self, arr):
        self.arr = np.ascontiguousarray(arr, dtype=np.float32)   # sitk arrays are indexed [z, y, x]


def GetImageFromArray(arr):
    return _Image(arr)


def WriteImage(image, path):
    a = image.arr
    nz, ny, nx = a.shape
    hdr = bytearray(348)
    struct.pack_into("<i", hdr, 0, 348)
    struct.pack_into("<8h", hdr, 40, 3, nx, ny, nz, 1, 1, 1, 1)      # dim
    struct.pack_into("<h", hdr, 70, 16)                               # datatype float32
    struct.pack_into("<h", hdr, 72, 32)                               # bitpix
    struct.pack_into("<8f", hdr, 76, 1.0, 1.0, 1.0, 1.0, 0.0, 0.0, 0.0, 0.0)   # pixdim
    struct.pack_into("<f", hdr, 108, 352.0)                           # vox_offset
    struct.pack_into("<f", hdr, 112, 1.0)                             # scl_slope
    hdr[344:348] = b"n+1\0"
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "wb") as f:
        f.write(bytes(hdr) + b"\0\0\0\0" + a.tobytes())
