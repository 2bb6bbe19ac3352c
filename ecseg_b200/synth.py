"""Seeded synthetic inputs for the metaseg hot path (no datasets are reachable offline).

* synth_dapi       -- a metaphase-spread-like DAPI image: dark noisy
                      background, a few large dim nuclei, 40-80 bright rotated-ellipse chromosomes
                      around a metaphase centre and 50-300 small bright ecDNA discs.
* synth_label_map  -- a 4-class label map at the `np.argmax -> meta_inference` boundary
                      (reference src/utils.py:118-119), SURVEY.md §8d config 4: blobs of every
                      class plus salt pixels that create 1-px components and holes.
* synth_fish       -- RGB image with DAPI in the blue channel and green/red FISH spots (config 5).

Only numpy + OpenCV drawing primitives with integer parameters are used, so a given seed gives
the same bytes on every machine with the same OpenCV build.
"""
from __future__ import annotations

import cv2
import numpy as np


def _rng(seed):
    return np.random.Generator(np.random.PCG64(int(seed)))


def synth_dapi(seed: int, h: int = 2048, w: int = 2048, dtype: str = "u8", rgb: bool = False,
               invert: bool = False) -> np.ndarray:
    rng = _rng(seed)
    img = np.zeros((h, w), np.float32)
    s = min(h, w) / 2048.0
    # nuclei: large dim filled ellipses
    for _ in range(int(rng.integers(2, 5))):
        cy, cx = int(rng.integers(0, h)), int(rng.integers(0, w))
        ax = (int(max(8, rng.integers(80, 200) * s)), int(max(8, rng.integers(80, 200) * s)))
        cv2.ellipse(img, (cx, cy), ax, float(rng.integers(0, 180)), 0, 360,
                    float(rng.integers(90, 140)), -1)
    # metaphase: chromosomes around a centre
    my, mx = int(rng.integers(h // 4, 3 * h // 4)), int(rng.integers(w // 4, 3 * w // 4))
    for _ in range(int(rng.integers(40, 81))):
        cy = int(np.clip(my + rng.normal(0, 150 * s), 0, h - 1))
        cx = int(np.clip(mx + rng.normal(0, 150 * s), 0, w - 1))
        ax = (int(max(3, rng.integers(15, 40) * max(s, 0.5))), int(max(2, rng.integers(5, 10) * max(s, 0.5))))
        cv2.ellipse(img, (cx, cy), ax, float(rng.integers(0, 180)), 0, 360,
                    float(rng.integers(170, 230)), -1)
    # ecDNA: small bright discs
    for _ in range(int(rng.integers(50, 301))):
        cy = int(np.clip(my + rng.normal(0, 300 * s), 0, h - 1))
        cx = int(np.clip(mx + rng.normal(0, 300 * s), 0, w - 1))
        cv2.circle(img, (cx, cy), int(rng.integers(2, 6)), float(rng.integers(150, 220)), -1)
    img = cv2.GaussianBlur(img, (0, 0), 1.5)
    img += np.clip(rng.normal(12, 4, (h, w)), 0, None).astype(np.float32)
    img += rng.normal(0, 1, (h, w)).astype(np.float32) * np.sqrt(np.maximum(img, 0)) * 0.3
    img = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    if invert:
        img = 255 - img
    if dtype == "u16":
        out = img.astype(np.uint16) * 257
        out += rng.integers(0, 128, out.shape).astype(np.uint16) * (out < 65000)
    else:
        out = img
    if rgb:
        planes = [np.zeros_like(out), np.zeros_like(out), out]
        out = np.stack(planes, axis=-1)  # RGB order, DAPI in channel 2 (reference image_tools.py:88-89)
    return out


def synth_label_map(seed: int, h: int = 2048, w: int = 2048, salt: int | None = None,
                    dtype=np.uint8) -> np.ndarray:
    """Random 4-class label map: 0 background, 1 nucleus, 2 chromosome, 3 ecDNA."""
    rng = _rng(seed)
    lab = np.zeros((h, w), np.uint8)
    s = min(h, w) / 2048.0
    n_nuc = int(rng.integers(0, 5))
    for _ in range(n_nuc):
        c = (int(rng.integers(0, w)), int(rng.integers(0, h)))
        ax = (int(max(4, rng.integers(60, 220) * s)), int(max(4, rng.integers(60, 220) * s)))
        cv2.ellipse(lab, c, ax, float(rng.integers(0, 180)), 0, 360, 1, -1)
    my, mx = int(rng.integers(h // 4, 3 * h // 4)), int(rng.integers(w // 4, 3 * w // 4))
    for _ in range(int(rng.integers(0, 90))):
        cy = int(np.clip(my + rng.normal(0, 140 * s), 0, h - 1))
        cx = int(np.clip(mx + rng.normal(0, 140 * s), 0, w - 1))
        ax = (int(max(2, rng.integers(10, 40) * max(s, 0.3))), int(max(1, rng.integers(3, 10) * max(s, 0.3))))
        cv2.ellipse(lab, (cx, cy), ax, float(rng.integers(0, 180)), 0, 360, 2, -1)
    for _ in range(int(rng.integers(0, 300))):
        cy = int(np.clip(my + rng.normal(0, 300 * s), 0, h - 1))
        cx = int(np.clip(mx + rng.normal(0, 300 * s), 0, w - 1))
        cv2.circle(lab, (cx, cy), int(rng.integers(1, 7)), 3, -1)
    # a few rings (holes to fill) of class 1 / 2
    for _ in range(int(rng.integers(0, 6))):
        c = (int(rng.integers(0, w)), int(rng.integers(0, h)))
        r = int(max(3, rng.integers(6, 60) * max(s, 0.3)))
        cv2.circle(lab, c, r, int(rng.integers(1, 3)), int(rng.integers(1, 4)))
    if salt is None:
        salt = int(2000 * s * s) + 8
    ys = rng.integers(0, h, salt)
    xs = rng.integers(0, w, salt)
    lab[ys, xs] = rng.integers(0, 4, salt).astype(np.uint8)
    return lab.astype(dtype)


def synth_noise_label_map(seed: int, h: int, w: int, p=(0.55, 0.15, 0.2, 0.1), block: int = 1) -> np.ndarray:
    """Pure-noise (optionally blocky) label map: the adversarial case for CCL."""
    rng = _rng(seed)
    hb, wb = -(-h // block), -(-w // block)
    lab = rng.choice(4, size=(hb, wb), p=p).astype(np.uint8)
    if block > 1:
        lab = np.kron(lab, np.ones((block, block), np.uint8))[:h, :w]
    return np.ascontiguousarray(lab)


def synth_fish(seed: int, h: int = 2048, w: int = 2048, dtype: str = "u8") -> np.ndarray:
    """RGB image: DAPI in B, green / red FISH spots in G / R (reference image_tools.py:136-146)."""
    rng = _rng(seed + 7919)
    dapi = synth_dapi(seed, h, w)
    planes = [np.zeros((h, w), np.float32), np.zeros((h, w), np.float32), dapi.astype(np.float32)]
    bright = np.argwhere(dapi > 120)
    for ch in (0, 1):
        n = int(rng.integers(20, 120))
        if len(bright):
            pick = bright[rng.integers(0, len(bright), n)]
        else:
            pick = np.stack([rng.integers(0, h, n), rng.integers(0, w, n)], 1)
        for (y, x) in pick:
            cv2.circle(planes[ch], (int(x), int(y)), int(rng.integers(1, 5)), float(rng.integers(100, 256)), -1)
        # a few larger signals (HSR-like, survive the 20-px small-object filter) and dim background speckle
        for (y, x) in pick[: max(1, n // 6)]:
            cv2.ellipse(planes[ch], (int(x), int(y)), (int(rng.integers(4, 9)), int(rng.integers(2, 5))),
                        float(rng.integers(0, 180)), 0, 360, float(rng.integers(100, 256)), -1)
        planes[ch] += rng.integers(0, 60, (h, w)).astype(np.float32)
    out = np.clip(np.rint(np.stack(planes, -1)), 0, 255).astype(np.uint8)
    if dtype == "u16":
        out = out.astype(np.uint16) * 257
    return out
