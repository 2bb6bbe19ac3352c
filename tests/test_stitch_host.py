"""CPU-only: the closed-form stitch ownership rule of ecseg_b200/csrc/stitch.cuh, executed on the host (tests/hostcheck/
stitch_host.cu), against the provenance codes of the reference's own patches2im_overlap (src/image_tools.py:188-252)
frozen in tests/golden/tiling.npz, and against the oracle on more shapes (single tile column / row, exact multiples of
the 206-pixel prediction window, the h_l == w_l case that leaves the right strip unwritten)."""
import ctypes
import hashlib
import os
import subprocess

import numpy as np
import pytest

from oracle import metaseg_oracle as mo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HC = os.path.join(ROOT, "tests", "hostcheck")


@pytest.fixture(scope="module")
def hc():
    subprocess.run(["make", "-C", HC, "libstitch_host.so"], check=True, capture_output=True)
    h = ctypes.CDLL(os.path.join(HC, "libstitch_host.so"))
    h.hostcheck_stitch_codes.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    return h


def codes(hc, h, w):
    out = np.zeros((h, w), np.int32)
    n = hc.hostcheck_stitch_codes(h, w, out.ctypes.data)
    assert n > 0
    return out, n


def oracle_codes(h, w):
    pos = mo.tile_positions(h, w)
    yy, xx = np.mgrid[0:256, 0:256]
    prov = np.zeros((len(pos), 256, 256, 4), np.float32)
    for k in range(len(pos)):
        prov[k, :, :, 0] = k * 65536 + yy * 256 + xx + 1
    return mo.patches2im_overlap(prov, pos)[:, :, 0].astype(np.int32), len(pos)


def test_closed_form_equals_reference_provenance(hc, golden):
    g = golden("tiling")
    for h, w in g["shapes"]:
        key = f"{h}x{w}"
        code, n = codes(hc, int(h), int(w))
        assert n == len(g["pos_" + key]), key
        assert int((code == 0).sum()) == int(g["nzero_" + key]), key
        sha = np.frombuffer(hashlib.sha256(code.tobytes()).digest(), np.uint8)
        assert np.array_equal(sha, g["sha_" + key]), key


@pytest.mark.parametrize("shape", [(256, 256), (256, 700), (700, 256), (462, 462), (462, 470), (463, 462), (668, 668),
                                   (300, 330), (1040, 1392), (257, 513), (1000, 1000)])
def test_closed_form_equals_oracle_on_more_shapes(hc, shape):
    code, n = codes(hc, *shape)
    want, n_ref = oracle_codes(*shape)
    assert n == n_ref
    assert np.array_equal(code, want), shape
