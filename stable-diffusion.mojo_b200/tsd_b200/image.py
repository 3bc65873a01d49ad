"""PNG output of the image pipeline.generate returns (pipeline.mojo:127-128: a (3,S,S) float matrix in
0..255; the reference never stores it).  The encoder runs in libtsd_b200.so (csrc/host_io.cu)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import TsdError


def encode_png(img) -> bytes:
    """(c,h,w) float image, c in {1,3,4}, values 0..255 -> PNG bytes (8 bit, round half up, clamped)."""
    a = np.ascontiguousarray(img, np.float32)
    if a.ndim != 3:
        raise TsdError(1, "png: image must be (c,h,w)")
    L = _lib.lib()
    c, h, w = a.shape
    size = C.c_int64()
    rc = L.tsd_png_encode(a.ctypes.data, c, h, w, None, 0, C.byref(size))
    if rc:
        raise TsdError(rc, "png: channels must be 1, 3 or 4 and the image non-empty")
    buf = (C.c_uint8 * size.value)()
    rc = L.tsd_png_encode(a.ctypes.data, c, h, w, buf, size.value, C.byref(size))
    if rc:
        raise TsdError(rc, "png: encode failed")
    return bytes(buf)


def save_png(path, img) -> None:
    a = np.ascontiguousarray(img, np.float32)
    if a.ndim != 3:
        raise TsdError(1, "png: image must be (c,h,w)")
    rc = _lib.lib().tsd_png_write(str(path).encode(), a.ctypes.data, *a.shape)
    if rc:
        raise TsdError(rc, f"png: cannot write {path}")
