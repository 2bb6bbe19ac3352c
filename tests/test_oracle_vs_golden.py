"""The CPU oracle (oracle/metaseg_oracle.py) against golden vectors produced by running the
reference's own functions (oracle/make_golden.py).  CPU only."""
import hashlib
import warnings

import numpy as np
import pytest

from oracle import metaseg_oracle as mo
from oracle.fake_model import FakeModel


def test_postprocess_full_pipeline(golden):
    g = golden("postproc")
    n = int(g["n_cases"])
    assert n >= 40
    for i in range(n):
        m = g[f"in_{i}"].astype(np.int64)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out = mo.meta_inference(m.copy())
        assert np.array_equal(out, g[f"out_{i}"]), f"case {i}"
        assert tuple(g[f"cnt_{i}"]) == mo.count_cc(out == 3), f"case {i}"
        assert tuple(g[f"cnt_in_{i}"]) == mo.count_cc(m == 3), f"case {i}"


def test_postprocess_single_steps(golden):
    g = golden("postproc")
    for i in range(int(g["n_cases"])):
        m = g[f"in_{i}"].astype(np.int64)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            assert np.array_equal(mo.fill_holes(m.copy(), 1), g[f"fill1_{i}"]), i
            assert np.array_equal(mo.fill_holes(m.copy(), 2), g[f"fill2_{i}"]), i
            assert np.array_equal(mo.size_thresh(m.copy()), g[f"size_{i}"]), i
            assert np.array_equal(mo.merge_comp(m.copy(), 1), g[f"merge1_{i}"]), i
            assert np.array_equal(mo.merge_comp(m.copy(), 2), g[f"merge2_{i}"]), i


def test_merge_comp_is_noop_inside_pipeline(golden):
    """after the ecDNA boundary erase no (class U ec) component is mixed."""
    g = golden("postproc")
    for i in range(int(g["n_cases"])):
        m = g[f"in_{i}"].astype(np.int64)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out = mo.meta_inference(m.copy(), with_merge=False)
        assert np.array_equal(out, g[f"out_{i}"]), f"case {i}"


def test_merge_comp_standalone_is_not_noop(golden):
    g = golden("postproc")
    changed = sum(not np.array_equal(g[f"merge1_{i}"], g[f"in_{i}"]) for i in range(int(g["n_cases"])))
    assert changed > 5


def test_tiling_positions_and_stitch(golden):
    g = golden("tiling")
    for h, w in g["shapes"]:
        key = f"{h}x{w}"
        pos = mo.tile_positions(int(h), int(w))
        assert np.array_equal(pos, g["pos_" + key]), key
        n = len(pos)
        yy, xx = np.mgrid[0:256, 0:256]
        prov = np.zeros((n, 256, 256, 4), np.float32)
        for k in range(n):
            prov[k, :, :, 0] = k * 65536 + yy * 256 + xx + 1
        code = mo.patches2im_overlap(prov, pos)[:, :, 0].astype(np.int32)
        assert int((code == 0).sum()) == int(g["nzero_" + key]), key
        sha = np.frombuffer(hashlib.sha256(code.tobytes()).digest(), np.uint8)
        assert np.array_equal(sha, g["sha_" + key]), key
        if "code_" + key in g.files:
            assert np.array_equal(code, g["code_" + key]), key
    assert int(g["nzero_2048x2048"]) == 44800       # the image_tools.py:242 strip (SURVEY finding 5)
    assert int(g["nzero_1040x1392"]) == 0


def test_preprocess(golden):
    g = golden("preprocess")
    for k in g["names"]:
        out = mo.meta_preprocess(g["in_" + str(k)].copy())
        assert out.dtype == np.uint8
        assert np.array_equal(out, g["out_" + str(k)]), k
    ramp = np.arange(65536, dtype=np.uint16).reshape(256, 256)
    assert np.array_equal(mo.u16_to_u8(ramp), g["u16_ramp_out"])


def test_otsu_matches_opencv():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    for s in range(30):
        a = rng.integers(0, 256, (64, 80)).astype(np.uint8)
        a[: s * 2] //= (s % 5) + 1
        t, _ = cv2.threshold(a, 0, 1, cv2.THRESH_BINARY + cv2.THRESH_OTSU)
        assert int(t) == mo.otsu_threshold_from_hist(np.bincount(a.ravel(), minlength=256))


def test_meta_segment_end_to_end(golden):
    g = golden("segment")
    for k in g["names"]:
        k = str(k)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            lab, pre = mo.meta_segment_array(FakeModel(), g["in_" + k].copy())
        assert np.array_equal(lab, g["lab_" + k]), k
        assert np.array_equal(255 - pre, g["dapi_" + k]), k
        assert tuple(g["cnt_" + k]) == mo.count_cc(lab == 3), k


def test_quantise_argmax_ties_and_rounding():
    p = np.zeros((1, 6, 4))
    p[0, 0] = [0.25, 0.25, 0.25, 0.25]            # tie -> class 0
    p[0, 1] = [0.1, 0.4500001, 0.4499999, 0.0]    # both round to 115 -> first max = 1
    p[0, 2] = [0.0, 0.0, 0.5 / 255, 1.5 / 255]    # 0.5 -> 0 (half-even), 1.5 -> 2
    p[0, 3] = [0.0, 2.5 / 255, 0.0, 2.4 / 255]    # 2.5 -> 2, 2.4 -> 2: tie -> class 1
    p[0, 4] = [0.0, 0.0, 0.0, 1.0]
    p[0, 5] = [0.0, 0.0, 0.0, 0.0]                # the unwritten strip -> class 0
    assert mo.quantise_argmax(p).tolist() == [[0, 1, 3, 1, 3, 0]]
    with pytest.raises(ValueError):
        mo.quantise_argmax(np.full((1, 1, 4), 1.5))


def test_overlay_palette():
    lab = np.array([[0, 1], [2, 3]])
    assert mo.overlay_rgba(lab).tolist() == [[[56, 108, 176, 255], [255, 255, 153, 255]],
                                             [[127, 201, 127, 255], [240, 2, 127, 255]]]


OVERLAY_KEYS = ("num_ecDNA", "num_FISH", "num_ecDNA_FISH", "num_HSR", "num_FISH2", "num_FISH_FISH2", "num_ecDNA_FISH2",
                "num_ecDNA_FISH_FISH2", "num_HSR2")


def _flat_overlay(res):
    out = []
    for k in OVERLAY_KEYS:
        v = res[k]
        out.extend(v if isinstance(v, tuple) else [v])
    return [int(x) for x in out]


def test_meta_overlay_counts(golden):
    """meta_overlay per-image body (reference src/meta_overlay.py:59-83) incl. the np.unique()[1:] quirk cases."""
    g = golden("overlay")
    for i in range(int(g["n_cases"])):
        res = mo.overlay_counts(g[f"img_{i}"], g[f"seg_{i}"], int(g[f"sens_{i}"]))
        assert _flat_overlay(res) == [int(v) for v in g[f"out_{i}"]], i
    assert mo.overlay_counts(g["img_0"][..., 0], g["seg_0"], 85) is None      # non-RGB images are skipped


def test_config4_full_size_maps(golden):
    """The oracle on full-size config-4 maps against the frozen reference run (a sample: ~1.5 s of CPU per map)."""
    from ecseg_b200 import synth
    g = golden("config4")
    for seed in (0, 17, 63):
        m = synth.synth_label_map(seed, 2048, 2048).astype(np.int64)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out = mo.meta_inference(m)
        assert mo.count_cc(out == 3) == tuple(int(v) for v in g["cnt"][seed]), seed
        assert hashlib.sha256(out.astype(np.uint8).tobytes()).digest() == g["sha"][seed].tobytes(), seed
