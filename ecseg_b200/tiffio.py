"""Image input of the metaseg driver: skimage.io.imread semantics (reference src/utils.py:110).

Uncompressed little-endian strip TIFFs (what microscopes and the synthetic generator write) are read by
libecseg_b200's ecseg_tiff_read straight into a caller-provided (pinned) buffer; every other flavour goes through
the general decoder of utils.imread (cv2), channel order RGB(A) either way."""
from __future__ import annotations

import ctypes
from ctypes import byref, c_int, c_void_p

import numpy as np

from . import _lib


def probe(path: str):
    """(h, w, ch, bytes_per_sample) when the fast reader handles `path`, else None."""
    lib = _lib.load()
    h, w, ch, bps = c_int(), c_int(), c_int(), c_int()
    rc = lib.ecseg_tiff_read(path.encode(), None, 0, byref(h), byref(w), byref(ch), byref(bps))
    if rc == -1:
        raise FileNotFoundError(path)
    return (h.value, w.value, ch.value, bps.value) if rc == 0 else None


def read_into(path: str, buf: np.ndarray):
    """Decode `path` into the front of the uint8 buffer `buf`; returns the image as an array VIEW of buf
    ([H,W] or [H,W,C], uint8 or uint16).  Raises ValueError when buf is too small."""
    lib = _lib.load()
    h, w, ch, bps = c_int(), c_int(), c_int(), c_int()
    rc = lib.ecseg_tiff_read(path.encode(), buf.ctypes.data_as(c_void_p), buf.size, byref(h), byref(w), byref(ch), byref(bps))
    if rc == -1:
        raise FileNotFoundError(path)
    if rc == -2:
        raise ValueError(f"{path}: image larger than the staging buffer")
    if rc == 0:
        n = h.value * w.value * ch.value * bps.value
        a = buf[:n].view(np.uint16 if bps.value == 2 else np.uint8)
        return a.reshape((h.value, w.value) if ch.value == 1 else (h.value, w.value, ch.value))
    from .utils import imread
    a = imread(path)
    if a.dtype not in (np.uint8, np.uint16) or a.ndim not in (2, 3):
        raise ValueError(f"{path}: unsupported sample type {a.dtype}")
    if a.nbytes > buf.size:
        raise ValueError(f"{path}: image larger than the staging buffer")
    v = buf[:a.nbytes].view(a.dtype).reshape(a.shape)
    np.copyto(v, a)
    return v


def imread(path: str) -> np.ndarray:
    """Stand-alone read (allocates): fast path when possible, else utils.imread."""
    info = probe(path) if path.lower().endswith(('.tif', '.tiff')) else None
    if info is None:
        from .utils import imread as general
        return general(path)
    h, w, ch, bps = info
    buf = np.empty(h * w * ch * bps, np.uint8)
    return read_into(path, buf)
