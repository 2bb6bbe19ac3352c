"""CPU-only: the host half of the artefact writers and the TIFF input reader.

* the scanline encoder of csrc/png_deflate.cuh, executed sequentially on the host by tests/hostcheck (same token,
  framing and Adler code the CUDA kernels run), must inflate with zlib to the PNG-filtered scanlines and decode with
  cv2 to exactly the oracle's overlay (plt.imsave with the ListedColormap, reference src/metaseg.py:47-52);
* ecseg_npy_header / ecseg_tiff_header / ecseg_png_wrap / ecseg_crc32 against numpy, cv2 and zlib;
* ecseg_tiff_read against the general decoder on every layout it accepts, and its refusals."""
import ctypes
import io
import os
import subprocess
import zlib

import cv2
import numpy as np
import pytest

from ecseg_b200 import _lib, tiffio
from oracle import metaseg_oracle as mo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HC = os.path.join(ROOT, "tests", "hostcheck")


@pytest.fixture(scope="module")
def lib():
    if not os.path.isfile(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return _lib.load()


@pytest.fixture(scope="module")
def hc():
    subprocess.run(["make", "-C", HC], check=True, capture_output=True)
    h = ctypes.CDLL(os.path.join(HC, "libpngdef_host.so"))
    h.hostcheck_zlib_stream.restype = ctypes.c_size_t
    h.hostcheck_zlib_stream.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t]
    h.hostcheck_zlib_cap.restype = ctypes.c_size_t
    h.hostcheck_zlib_cap.argtypes = [ctypes.c_int, ctypes.c_int]
    h.hostcheck_tiff_parse.argtypes = [ctypes.c_void_p, ctypes.c_size_t] + [ctypes.POINTER(ctypes.c_int)] * 4
    return h


def png_filter_reference(rgba):
    """PNG scanlines with filter 1 (Sub) on row 0 and filter 2 (Up) below, as the encoder chooses them."""
    h, w, _ = rgba.shape
    raw = rgba.reshape(h, w * 4).astype(np.int16)
    out = np.zeros((h, w * 4 + 1), np.uint8)
    out[0, 0] = 1
    left = np.concatenate([np.zeros(4, np.int16), raw[0, :-4]])
    out[0, 1:] = ((raw[0] - left) & 0xFF).astype(np.uint8)
    out[1:, 0] = 2
    out[1:, 1:] = ((raw[1:] - raw[:-1]) & 0xFF).astype(np.uint8)
    return out.tobytes()


def encode_host(hc, lib, lab):
    lab = np.ascontiguousarray(lab, np.uint8)
    h, w = lab.shape
    cap = hc.hostcheck_zlib_cap(h, w)
    buf = np.zeros(cap + 64, np.uint8)
    n = hc.hostcheck_zlib_stream(lab.ctypes.data, h, w, buf.ctypes.data + 41, cap)
    assert n > 0
    total = lib.ecseg_png_wrap(buf.ctypes.data, n, h, w)
    return buf[:total], buf[41:41 + n]


def label_maps():
    rng = np.random.default_rng(7)
    for shape in [(1, 1), (1, 7), (2, 8), (5, 9), (16, 33), (64, 64), (131, 517), (260, 1031)]:
        yield rng.integers(0, 4, shape)                 # incompressible: every pixel a literal run
        yield np.zeros(shape, np.int64)                 # one colour: maximal zero runs (258-byte matches)
        yield np.full(shape, 3)
        sparse = np.zeros(shape, np.int64)
        m = rng.random(shape) < 0.03
        sparse[m] = rng.integers(1, 4, int(m.sum()))
        yield sparse
    g = np.load(os.path.join(ROOT, "tests", "golden", "postproc.npz"))
    yield g["out_0"]
    yield g["in_1"]


def test_scanline_encoder_inflates_to_filtered_overlay(hc, lib):
    for lab in label_maps():
        png, z = encode_host(hc, lib, lab)
        rgba = mo.overlay_rgba(np.asarray(lab))
        assert zlib.decompress(z.tobytes()) == png_filter_reference(rgba), lab.shape   # deflate + adler32 are valid
        img = cv2.imdecode(png, cv2.IMREAD_UNCHANGED)                                  # chunk framing + CRCs are valid
        assert img is not None and img.shape == rgba.shape
        assert np.array_equal(img[..., [2, 1, 0, 3]], rgba), lab.shape


def test_zero_run_lengths_cover_every_match_length(hc, lib):
    # a single non-background pixel at column x leaves zero runs of every length modulo 258 on both sides
    for w in (66, 130, 259, 300):
        lab = np.zeros((3, w), np.uint8)
        for x in range(w):
            lab[:] = 0
            lab[1, x] = 2
            png, _ = encode_host(hc, lib, lab)
            img = cv2.imdecode(png, cv2.IMREAD_UNCHANGED)
            assert np.array_equal(img[..., [2, 1, 0, 3]], mo.overlay_rgba(lab)), (w, x)


def test_png_size_bound_and_typical_ratio(hc, lib):
    from ecseg_b200 import synth  # noqa: F401  (shape generator only)
    rng = np.random.default_rng(3)
    noise = rng.integers(0, 4, (64, 200))
    png, z = encode_host(hc, lib, noise)
    assert z.size <= hc.hostcheck_zlib_cap(64, 200)
    yy, xx = np.mgrid[:512, :512]
    lab = np.zeros((512, 512), np.uint8)
    for _ in range(20):
        cy, cx, r = rng.integers(0, 512, 2).tolist() + [int(rng.integers(3, 60))]
        lab[(yy - cy) ** 2 + (xx - cx) ** 2 < r * r] = rng.integers(1, 4)
    png, _ = encode_host(hc, lib, lab)
    assert png.size < 0.03 * lab.size * 4      # blobs on a flat background: a few % of the raw RGBA


def test_crc32_matches_zlib(lib):
    rng = np.random.default_rng(0)
    for n in (0, 1, 7, 8, 9, 63, 1000, 65537):
        a = rng.integers(0, 256, n, dtype=np.uint8)
        assert lib.ecseg_crc32(0, a.ctypes.data, n) == zlib.crc32(a.tobytes())
    a = rng.integers(0, 256, 100, dtype=np.uint8)
    part = lib.ecseg_crc32(0, a.ctypes.data, 40)
    assert lib.ecseg_crc32(part, a.ctypes.data + 40, 60) == zlib.crc32(a.tobytes())


def test_npy_header_is_what_np_save_writes(lib):
    for h, w in [(1, 1), (256, 256), (1040, 1392), (2048, 2048), (12345, 7)]:
        want = io.BytesIO()
        np.save(want, np.zeros((h, w), np.int64)[:0])           # header only differs by the shape text
        buf = np.zeros(4096, np.uint8)
        n = lib.ecseg_npy_header(buf.ctypes.data, buf.size, h, w)
        assert n % 64 == 0 and lib.ecseg_npy_header(None, 0, h, w) == n
        payload = np.arange(h * w, dtype=np.int64).reshape(h, w) % 4 if h * w < 1 << 22 else None
        if payload is not None:
            ref = io.BytesIO()
            np.save(ref, payload)
            assert buf[:n].tobytes() + payload.tobytes() == ref.getvalue(), (h, w)    # byte-identical to np.save
            back = np.load(io.BytesIO(buf[:n].tobytes() + payload.tobytes()))
            assert back.dtype == np.int64 and np.array_equal(back, payload)


def test_tiff_header_decodes(lib):
    rng = np.random.default_rng(1)
    for h, w in [(1, 1), (300, 517), (256, 256)]:
        plane = rng.integers(0, 256, (h, w), dtype=np.uint8)
        buf = np.zeros(128 + h * w, np.uint8)
        assert lib.ecseg_tiff_header(buf.ctypes.data, h, w) == 128
        buf[128:] = plane.ravel()
        img = cv2.imdecode(buf, cv2.IMREAD_UNCHANGED)
        assert img is not None and img.dtype == np.uint8 and np.array_equal(img, plane)
        assert tiffio_probe_bytes(buf.tobytes(), lib) == (h, w, 1, 1)      # and our own reader takes our own files


def tiffio_probe_bytes(data, lib, tmp=[None]):
    import tempfile
    with tempfile.NamedTemporaryFile(suffix=".tif") as f:
        f.write(data)
        f.flush()
        return tiffio.probe(f.name)


def test_tiff_read_matches_general_decoder(lib, tmp_path):
    from ecseg_b200.utils import imread as general
    rng = np.random.default_rng(5)
    cases = {
        "g8": rng.integers(0, 256, (300, 517), dtype=np.uint8),
        "g16": rng.integers(0, 65536, (257, 300), dtype=np.uint16),
        "rgb8": rng.integers(0, 256, (280, 290, 3), dtype=np.uint8),
        "rgb16": rng.integers(0, 65536, (260, 333, 3), dtype=np.uint16),
        "rgba8": rng.integers(0, 256, (256, 256, 4), dtype=np.uint8),
    }
    for name, a in cases.items():
        p = str(tmp_path / (name + ".tif"))
        bgr = a if a.ndim == 2 else (a[..., ::-1] if a.shape[2] == 3 else a[..., [2, 1, 0, 3]])
        assert cv2.imwrite(p, np.ascontiguousarray(bgr), [cv2.IMWRITE_TIFF_COMPRESSION, 1])
        info = tiffio.probe(p)
        assert info == (a.shape[0], a.shape[1], 1 if a.ndim == 2 else a.shape[2], a.dtype.itemsize), (name, info)
        got = tiffio.imread(p)
        assert got.dtype == a.dtype and np.array_equal(got, a), name            # stored (RGB) order, like skimage
        assert np.array_equal(got, general(p)), name
        small = np.empty(10, np.uint8)
        with pytest.raises(ValueError):
            tiffio.read_into(p, small)
    # LZW (cv2's default) is not handled by the fast reader: probe says so, read_into falls back and still agrees
    p = str(tmp_path / "lzw.tif")
    cv2.imwrite(p, cases["g8"])
    assert tiffio.probe(p) is None
    buf = np.empty(cases["g8"].nbytes, np.uint8)
    assert np.array_equal(tiffio.read_into(p, buf), cases["g8"])
    with pytest.raises(FileNotFoundError):
        tiffio.probe(str(tmp_path / "missing.tif"))


def test_tiff_parser_refusals(hc):
    def parse(b):
        a = np.frombuffer(b, np.uint8)
        v = [ctypes.c_int() for _ in range(4)]
        return hc.hostcheck_tiff_parse(a.ctypes.data, a.size, *[ctypes.byref(x) for x in v])
    assert parse(b"MM\x00\x2a\x00\x00\x00\x08" + b"\0" * 64) == 1          # big endian
    assert parse(b"II\x2b\x00" + b"\0" * 64) == 1                          # BigTIFF
    assert parse(b"II\x2a\x00\xff\xff\x00\x00") == 2                       # IFD beyond the file
    assert parse(b"garbage!") == 1
