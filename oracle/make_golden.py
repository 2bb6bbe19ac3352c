#!/usr/bin/env python3
"""Generate tests/golden/*.npz by RUNNING THE REFERENCE'S OWN FUNCTIONS (TEST INFRASTRUCTURE).

Run in the development container only (needs /root/reference):

    python oracle/make_golden.py

/root/reference/src/{image_tools,utils}.py are imported unmodified through oracle/ref_harness
(tensorflow / skimage / matplotlib replaced by scipy/OpenCV-backed stubs).
The nested helpers of meta_inference (merge_comp, fill_holes, size_thresh) are lifted out of the
reference function's code object -- the bytecode that runs is the reference's.

Fixtures (all small, compressed):
  postproc.npz    label maps -> reference meta_inference result, count_cc tuple, and the three
                  nested helpers run standalone
  tiling.npz      im2patches_overlap positions + provenance-coded patches2im_overlap canvases
  preprocess.npz  meta_preprocess on u8/u16/gray/RGB/bright-background inputs
  segment.npz     utils.meta_segment end to end with a deterministic fake model (tif on disk)
  overlay.npz     meta_overlay's per-image counts (count_cc / count_colocalization / count_HSR)
  example.npz     BASELINE config 1: utils.meta_segment on input.tif := 255 - example_ecSeg/dapi.jpeg
                  (the only fixture the reference ships; SURVEY.md finding 0.4) with the fp32 torch-CPU
                  U-Net oracle standing in for Keras, seed-0 weights (`python oracle/make_golden.py example`)
  config4.npz     BASELINE config 4: reference meta_inference + count_cc on 64 synthetic 2048x2048 label maps
                  (count tuple, class histogram, SHA-256 of the final map) (`python oracle/make_golden.py config4`)
"""
from __future__ import annotations

import hashlib
import os
import sys
import tempfile
import types
import warnings

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from ecseg_b200 import synth  # noqa: E402
from oracle import ref_harness  # noqa: E402
from oracle.fake_model import FakeModel  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def lift_nested(fn):
    out = {}
    for c in fn.__code__.co_consts:
        if isinstance(c, types.CodeType):
            assert not c.co_freevars, c.co_name
            out[c.co_name] = types.FunctionType(c, fn.__globals__, c.co_name)
    return out


def ring_case(h=220, w=220):
    """Nucleus surrounded by a ring of chromosomes: the nucleus-in-metaphase rule
    (image_tools.py:72-81) fires."""
    lab = np.zeros((h, w), np.uint8)
    cv2.circle(lab, (w // 2, h // 2), 16, 1, -1)
    for k in range(28):
        a = 2 * np.pi * k / 28
        r = 34 + 14 * (k % 2)
        cv2.circle(lab, (int(w / 2 + r * np.cos(a)), int(h / 2 + r * np.sin(a))), 4, 2, -1)
    cv2.circle(lab, (30, 30), 14, 1, -1)      # a second nucleus far from the ring: survives
    cv2.circle(lab, (190, 40), 3, 3, -1)
    return lab


def postproc_cases():
    cases = []
    for s in range(24):
        h, w = [(96, 128), (130, 100), (160, 160), (200, 256)][s % 4]
        cases.append(synth.synth_label_map(s, h, w))
    for s in range(12):
        cases.append(synth.synth_noise_label_map(100 + s, 64 + 8 * s, 80, block=1 + s % 4))
    for s in range(6):   # empty-class maps: NaN paths of size_thresh
        m = synth.synth_label_map(200 + s, 120, 120)
        m[m == (s % 3) + 1] = 0
        cases.append(m)
    cases.append(np.zeros((40, 50), np.uint8))
    cases.append(np.full((40, 50), 3, np.uint8))
    cases.append(np.full((33, 47), 1, np.uint8))
    cases.append(ring_case())
    # ecDNA touching image borders / chromosomes, holes linked diagonally to the border
    m = np.zeros((64, 64), np.uint8)
    m[0:6, 0:9] = 3; m[20:40, 20:40] = 2; m[25:30, 25:30] = 0; m[38:46, 38:50] = 3
    m[50:64, 50:64] = 1; m[55, 55] = 0; m[62, 63] = 0; m[63, 62] = 0; m[10:14, 60:64] = 3
    cases.append(m)
    return cases


def gen_overlay(it):
    """meta_overlay per-image body (src/meta_overlay.py:59-83) through the reference's own helpers."""
    d = {}
    cases = []
    for s in range(10):
        h, w = [(160, 200), (256, 256), (300, 330)][s % 3]
        I = synth.synth_fish(s, h, w, dtype="u16" if s % 4 == 3 else "u8")
        seg = synth.synth_label_map(300 + s, h, w)
        cases.append((I, seg, [85, 10, 200, 0, 255][s % 5]))
    # quirk cases: everything ecDNA / everything chromosome (np.unique(...)[1:] drops the only component)
    I = synth.synth_fish(20, 64, 64)
    cases.append((I, np.full((64, 64), 3, np.uint8), 85))
    cases.append((I, np.full((64, 64), 2, np.uint8), 20))
    cases.append((np.full((64, 64, 3), 255, np.uint8), synth.synth_label_map(321, 64, 64), 85))
    for i, (I, seg, sens) in enumerate(cases):
        red = it.u16_to_u8(I)[..., 0] > sens
        green = it.u16_to_u8(I)[..., 1] > sens
        seg64 = seg.astype(np.int64)
        nuclei, chrom, ec = seg64 == 1, seg64 == 2, seg64 == 3
        fish = green * ~nuclei
        fish2 = red * ~nuclei
        vals = [it.count_cc(ec), it.count_cc(fish * ~chrom), it.count_colocalization(ec, fish),
                it.count_HSR(chrom, fish, 20), it.count_cc(fish2 * ~chrom),
                it.count_colocalization(fish * ~chrom, fish2 * ~chrom), it.count_colocalization(ec, fish2),
                it.count_colocalization(ec, fish2 * fish), it.count_HSR(chrom, fish2, 20)]
        flat = []
        for v in vals:
            flat.extend([int(x) for x in v] if isinstance(v, tuple) else [int(v)])
        d[f"img_{i}"] = I
        d[f"seg_{i}"] = seg.astype(np.uint8)
        d[f"sens_{i}"] = np.array(sens)
        # [n_ec, px_ec, n_fish, px_fish, n_ec_fish, n_hsr, n_fish2, px_fish2, n_fish_fish2, n_ec_fish2, n_ec_fish_fish2, n_hsr2]
        d[f"out_{i}"] = np.array(flat, np.int64)
    d["n_cases"] = np.array(len(cases))
    np.savez_compressed(os.path.join(OUT, "overlay.npz"), **d)
    print("overlay.npz:", len(cases), "cases")


def gen_config4(it, n_maps=64):
    """BASELINE config 4 at full size: the reference's meta_inference + count_cc (src/image_tools.py:15-84,114-119) on
    the first `n_maps` synthetic 2048x2048 label maps the bench cycles (ecseg_b200.synth.synth_label_map, seeds
    0..n_maps-1).  Frozen per map: the count_cc tuple, the class histogram and a SHA-256 of the final uint8 map."""
    d = {"cnt": np.zeros((n_maps, 2), np.int64), "hist": np.zeros((n_maps, 4), np.int64), "sha": np.zeros((n_maps, 32), np.uint8)}
    for s in range(n_maps):
        m = synth.synth_label_map(s, 2048, 2048)
        r = it.meta_inference(m.astype(np.int64).copy())
        n, px = it.count_cc(r == 3)
        d["cnt"][s] = (int(n), int(px))
        d["hist"][s] = np.bincount(r.ravel(), minlength=4)
        d["sha"][s] = np.frombuffer(hashlib.sha256(r.astype(np.uint8).tobytes()).digest(), np.uint8)
        if s % 8 == 0:
            print("config4 map", s, d["cnt"][s], flush=True)
    np.savez_compressed(os.path.join(OUT, "config4.npz"), **d)
    print("config4.npz:", n_maps, "maps")


EXAMPLE_JPEG = "/root/reference/example_ecSeg/dapi.jpeg"


def gen_example(it, ut):
    """BASELINE.json config 1.  example_ecSeg/input.tif is missing from the reference checkout
    (.MISSING_LARGE_BLOBS); dapi.jpeg is the reference's own dapi/ output for it, i.e. 255 - the
    pre-processed image (src/utils.py:112), so input.tif := 255 - dapi.jpeg (JPEG-lossy).  The reference's
    utils.meta_segment (src/utils.py:109-120) runs UNMODIFIED on that file; the model object is the fp32
    torch-CPU U-Net oracle (seed-0 weights) wrapped to record what predict_on_batch returned, and the raw
    (pre-meta_inference) label map is rebuilt from those predictions with the reference's own
    patches2im_overlap + img_as_ubyte + np.argmax (src/utils.py:116-118)."""
    from skimage import img_as_ubyte          # the harness stub the reference itself imports

    from ecseg_b200 import weights as wmod
    from oracle.unet_oracle import UNetOracle

    jpg = cv2.imread(EXAMPLE_JPEG, cv2.IMREAD_GRAYSCALE)
    assert jpg is not None and jpg.shape == (1040, 1392), "example_ecSeg/dapi.jpeg not found / unexpected shape"
    inp = (255 - jpg).astype(np.uint8)

    class Recorder:
        def __init__(self, net):
            self.net, self.tiles, self.preds = net, None, None

        def predict_on_batch(self, x):
            self.tiles = np.asarray(x)
            self.preds = self.net.predict_on_batch(self.tiles)
            return self.preds

    rec = Recorder(UNetOracle(wmod.make_weights(0), batch=5))
    with tempfile.TemporaryDirectory() as tmp:
        os.mkdir(os.path.join(tmp, "dapi"))
        p = os.path.join(tmp, "input.tif")
        cv2.imwrite(p, inp)
        final = ut.meta_segment(rec, p)
        dapi = cv2.imread(os.path.join(tmp, "dapi", "input.tif"), cv2.IMREAD_UNCHANGED)
    assert rec.tiles.shape == (35, 256, 256, 1) and rec.tiles.dtype == np.uint8
    _img, _patches, pos = it.im2patches_overlap(it.meta_preprocess(inp.copy())[..., None])
    canvas = it.patches2im_overlap(list(rec.preds), pos)
    q = img_as_ubyte(canvas)
    raw = np.argmax(q, axis=2)
    again = it.meta_inference(raw.copy())
    assert np.array_equal(again, final), "rebuilt raw label map does not reproduce meta_segment's result"
    srt = np.sort(q.astype(np.int16), axis=2)
    notie = srt[..., 3] != srt[..., 2]
    z = rec.net.predict_logits(rec.tiles)
    d = {
        "input": inp, "dapi": dapi, "pos": np.array(pos, np.int64),
        "raw": raw.astype(np.uint8),                     # label map at the src/utils.py:118 -> :119 boundary
        "notie": np.packbits(notie),                     # pixels whose quantised top-2 differ (agreement metric)
        "final": final.astype(np.uint8),                 # meta_segment's return value
        "count": np.array(it.count_cc(final == 3), np.int64),
        # a sparse, frozen view of the oracle's own arithmetic (every 8th pixel of every tile): lets the GPU box
        # check that ITS torch-CPU build reproduces the oracle that generated this file
        "logits_sub8": z[:, ::8, ::8, :].astype(np.float32),
        "logits_absmax": np.array(float(np.abs(z).max())),
        "hist_raw": np.bincount(raw.ravel(), minlength=4).astype(np.int64),
        "hist_final": np.bincount(final.ravel().astype(np.int64), minlength=4).astype(np.int64),
    }
    np.savez_compressed(os.path.join(OUT, "example.npz"), **d)
    print("example.npz: 1040x1392, 35 tiles; raw hist", d["hist_raw"], "final hist", d["hist_final"], "count", d["count"],
          "ties %.3f%%" % (100 * (1 - notie.mean())))


def main():
    os.makedirs(OUT, exist_ok=True)
    warnings.simplefilter("ignore")
    it, ut = ref_harness.load_reference()
    if len(sys.argv) > 1 and sys.argv[1] == "overlay":
        gen_overlay(it)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "example":
        gen_example(it, ut)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "config4":
        gen_config4(it)
        return
    nested = lift_nested(it.meta_inference)
    assert set(nested) >= {"merge_comp", "fill_holes", "size_thresh"}, nested.keys()

    # ---------------- post-processing ----------------
    d = {}
    cases = postproc_cases()
    fired = 0
    for i, m in enumerate(cases):
        d[f"in_{i}"] = m.astype(np.uint8)
        r = it.meta_inference(m.astype(np.int64).copy())
        d[f"out_{i}"] = r.astype(np.uint8)
        n, px = it.count_cc(r == 3)
        d[f"cnt_{i}"] = np.array([int(n), int(px)], np.int64)
        d[f"fill1_{i}"] = nested["fill_holes"](m.astype(np.int64).copy(), 1).astype(np.uint8)
        d[f"fill2_{i}"] = nested["fill_holes"](m.astype(np.int64).copy(), 2).astype(np.uint8)
        d[f"size_{i}"] = nested["size_thresh"](m.astype(np.int64).copy()).astype(np.uint8)
        d[f"merge1_{i}"] = nested["merge_comp"](m.astype(np.int64).copy(), 1).astype(np.uint8)
        d[f"merge2_{i}"] = nested["merge_comp"](m.astype(np.int64).copy(), 2).astype(np.uint8)
        n0, px0 = it.count_cc(m == 3)
        d[f"cnt_in_{i}"] = np.array([int(n0), int(px0)], np.int64)
    d["n_cases"] = np.array(len(cases))
    ring = ring_case()
    rr = it.meta_inference(ring.astype(np.int64).copy())
    assert (ring == 1).sum() > (rr == 1).sum() > 0, "ring case must remove exactly one nucleus"
    np.savez_compressed(os.path.join(OUT, "postproc.npz"), **d)
    print("postproc.npz:", len(cases), "cases")

    # ---------------- tiling / stitching ----------------
    d = {}
    shapes = [(256, 256), (256, 300), (300, 256), (256, 700), (600, 256), (300, 300), (300, 420), (462, 470), (520, 462), (700, 512), (462, 462),
              (1040, 1392), (2048, 2048), (2048, 2049)]
    for (h, w) in shapes:
        img = np.zeros((h, w, 1), np.uint8)
        _img, patches, pos = it.im2patches_overlap(img)
        pos = np.array(pos, np.int64)
        n = len(pos)
        prov = np.zeros((n, 256, 256, 4), np.float32)
        yy, xx = np.mgrid[0:256, 0:256]
        for k in range(n):
            prov[k, :, :, 0] = k * 65536 + yy * 256 + xx + 1     # +1: 0 means "never written"
        canvas = it.patches2im_overlap(list(prov), [list(p) for p in pos])
        code = canvas[:, :, 0].astype(np.int64)
        key = f"{h}x{w}"
        d["pos_" + key] = pos
        d["nzero_" + key] = np.array(int((code == 0).sum()))
        d["sha_" + key] = np.frombuffer(hashlib.sha256(code.astype(np.int32).tobytes()).digest(), np.uint8)
        if h * w <= 800 * 600:
            d["code_" + key] = code.astype(np.int32)
    d["shapes"] = np.array(shapes, np.int64)
    np.savez_compressed(os.path.join(OUT, "tiling.npz"), **d)
    print("tiling.npz:", len(shapes), "shapes")

    # ---------------- pre-processing ----------------
    d = {}
    pre_in = {
        "u8_gray": synth.synth_dapi(3, 300, 340),
        "u8_bright_bg": synth.synth_dapi(4, 280, 300, invert=True),
        "u16_gray": synth.synth_dapi(5, 300, 320, dtype="u16"),
        "u8_rgb": synth.synth_dapi(6, 290, 310, rgb=True),
        "u16_rgb_bright": synth.synth_dapi(7, 300, 300, dtype="u16", rgb=True, invert=True),
        "flat": np.full((260, 270), 17, np.uint8),
        "two_level": np.where(np.arange(300 * 300).reshape(300, 300) % 7 < 4, 200, 10).astype(np.uint8),
    }
    for k, v in pre_in.items():
        d["in_" + k] = v
        d["out_" + k] = it.meta_preprocess(v.copy())
    d["names"] = np.array(list(pre_in))
    allu16 = np.arange(65536, dtype=np.uint16).reshape(256, 256)
    d["u16_ramp_out"] = it.u16_to_u8(allu16)
    np.savez_compressed(os.path.join(OUT, "preprocess.npz"), **d)
    print("preprocess.npz:", len(pre_in), "inputs")

    # ---------------- meta_segment end to end with a fake model ----------------
    d = {}
    with tempfile.TemporaryDirectory() as tmp:
        os.mkdir(os.path.join(tmp, "dapi"))
        seg_in = {"a": synth.synth_dapi(11, 300, 330), "b": synth.synth_dapi(12, 462, 462, invert=True),
                  "c": synth.synth_dapi(13, 300, 300, dtype="u16", rgb=True)}
        for k, v in seg_in.items():
            p = os.path.join(tmp, k + ".tif")
            cv2.imwrite(p, v[..., ::-1] if v.ndim == 3 else v)
            lab = ut.meta_segment(FakeModel(), p)
            d["in_" + k] = v
            d["lab_" + k] = lab.astype(np.uint8)
            d["dapi_" + k] = cv2.imread(os.path.join(tmp, "dapi", k + ".tif"), cv2.IMREAD_UNCHANGED)
            d["cnt_" + k] = np.array(it.count_cc(lab == 3), np.int64)
    d["names"] = np.array(list(seg_in))
    np.savez_compressed(os.path.join(OUT, "segment.npz"), **d)
    print("segment.npz:", len(seg_in), "images")

    gen_overlay(it)
    gen_example(it, ut)
    gen_config4(it)


if __name__ == "__main__":
    main()
