"""CPU-only: the ownership-aware block skipping of the level-0 decoder chain (unet.cu owned_blocks / work lists) against
the ORACLE's stitch ownership map (oracle.metaseg_oracle.patches2im_overlap, pinned on the reference's
image_tools.py:188-252 by tests/golden/tiling.npz).  For every tile the whole-image path may only skip a block if no
pixel the stitcher takes from that tile lies within the layer's dependency margin of it:
   head: owned pixels   conv1-4: owned +-1   conv1-3: owned +-2   up1: owned +-3 (its blocks are 32 x 16 output pixels)."""
import ctypes
from ctypes import byref, c_int, c_void_p

import numpy as np
import pytest

from ecseg_b200 import _lib, spec
from ecseg_b200.engine import unet_work
from oracle import metaseg_oracle as mo

SHAPES = [(256, 256), (300, 330), (462, 470), (1040, 1392), (2048, 2048), (2048, 2049), (700, 512), (256, 700)]
LAYERS = {22: (0, 16, 16), 21: (1, 16, 16), 20: (2, 16, 16), 19: (3, 32, 16)}     # margin, block rows / cols in OUTPUT pixels
LEVEL1 = {18: (0, 16, 16), 17: (1, 16, 16), 16: (2, 32, 16)}                      # margin around R1 on the 128-px grid


def grow(mask, margin):
    """Every pixel within `margin` (Chebyshev) of a set pixel."""
    if not margin:
        return mask
    n = mask.shape[0]
    pad = np.pad(mask, margin)
    out = np.zeros_like(mask)
    for dy in range(2 * margin + 1):
        for dx in range(2 * margin + 1):
            out |= pad[dy:dy + n, dx:dx + n]
    return out


def owner_map(h, w):
    """[H,W] int: index of the tile whose prediction the stitcher writes to each pixel (-1: never written)."""
    pos = mo.tile_positions(h, w)
    n = len(pos)
    prov = np.zeros((n, 256, 256, 4), np.float32)
    for k in range(n):
        prov[k, :, :, 0] = k + 1
    canvas = mo.patches2im_overlap(prov, pos)
    return canvas[:, :, 0].astype(np.int64) - 1, pos


def mask_of(h, w, layer):
    lib = _lib.load()
    r, c = c_int(), c_int()
    assert lib.ecseg_debug_owned_blocks(h, w, layer, None, byref(r), byref(c)) == 0
    n = len(mo.tile_positions(h, w))
    m = np.zeros((n, r.value, c.value), np.uint8)
    assert lib.ecseg_debug_owned_blocks(h, w, layer, m.ctypes.data_as(c_void_p), byref(r), byref(c)) == 0
    return m


@pytest.mark.parametrize("shape", SHAPES)
def test_no_needed_pixel_is_skipped(shape):
    h, w = shape
    own, pos = owner_map(h, w)
    for layer, (margin, br, bc) in LAYERS.items():
        m = mask_of(h, w, layer)
        assert m.shape[1:] == (256 // br, 256 // bc)
        for t, (r0, c0) in enumerate(pos):
            mine = own[r0:r0 + 256, c0:c0 + 256] == t                    # pixels the stitcher takes from tile t
            if margin:                                                    # everything within `margin` px (Chebyshev)
                pad = np.pad(mine, margin)
                need = np.zeros_like(mine)
                for dy in range(2 * margin + 1):
                    for dx in range(2 * margin + 1):
                        need |= pad[dy:dy + 256, dx:dx + 256]
            else:
                need = mine
            blocks = need.reshape(256 // br, br, 256 // bc, bc).any(axis=(1, 3))
            assert not (blocks & (m[t] == 0)).any(), (shape, layer, t)


@pytest.mark.parametrize("shape", SHAPES)
def test_level1_blocks_cover_what_up1_reads(shape):
    """Decoder level 1 (128-px grid).  up1 is a stride-2 transposed conv, out[2i + k] += in[i] * K[k] (k = 0..2, cropped
    to 256): for its output to be valid on owned +-3 it reads conv2-4 at R1 = { i : {2i, 2i+1, 2i+2} meets owned +-3 },
    conv2-4 reads conv2-3 on R1 +-1, conv2-3 reads up2 on R1 +-2."""
    h, w = shape
    own, pos = owner_map(h, w)
    masks = {li: mask_of(h, w, li) for li in LEVEL1}
    for t, (r0, c0) in enumerate(pos):
        need0 = grow(own[r0:r0 + 256, c0:c0 + 256] == t, 3)          # up1's output, 256-px grid
        r1 = np.zeros((128, 128), bool)
        ys, xs = np.nonzero(need0)
        for ky in range(3):
            for kx in range(3):
                yy, xx = ys - ky, xs - kx
                ok = (yy >= 0) & (xx >= 0) & (yy % 2 == 0) & (xx % 2 == 0)
                r1[yy[ok] // 2, xx[ok] // 2] = True
        for li, (margin, br, bc) in LEVEL1.items():
            need = grow(r1, margin)
            blocks = need.reshape(128 // br, br, 128 // bc, bc).any(axis=(1, 3))
            m = masks[li][t]
            assert m.shape == blocks.shape
            assert not (blocks & (m == 0)).any(), (shape, li, t)


def test_skipping_is_substantial_but_bounded():
    ref, ex = unet_work(2048, 2048, labels_only=True)
    assert abs(ref - spec.unet_flops_per_tile() * 100) < 1e-6 * ref
    assert 0.95 < ex / ref < 0.97                      # 23 % of the chain, 4.1 % of the network
    ref0, ex0 = unet_work(2048, 2048, labels_only=False)
    assert ref0 == ex0 == ref                          # staged calls compute everything
    r1, e1 = unet_work(256, 256)
    assert r1 == e1                                    # a single tile owns all of itself
    m = mask_of(2048, 2048, 21)
    interior = m[5 * 10 + 5]                           # an interior tile: rows / columns 1..14 of 16
    assert interior[1:15, 1:15].all() and not interior[0].any() and not interior[:, 15].any()
