"""Oracle vs the LIVE reference (dev container only; auto-skips where /root/reference is absent, e.g.
on the GPU box).  Imports /root/reference/src/{image_tools,utils}.py unmodified through
oracle/ref_harness and compares the oracle's restatement on fresh seeded inputs that are NOT in the
frozen golden fixtures -- the pin that keeps oracle/metaseg_oracle.py honest."""
import os
import sys
import warnings

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from ecseg_b200 import synth  # noqa: E402
from oracle import metaseg_oracle as mo, ref_harness  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_harness.available(), reason="/root/reference not present")


@pytest.fixture(scope="module")
def ref():
    warnings.simplefilter("ignore")
    return ref_harness.load_reference()


def test_meta_inference_and_count_on_fresh_maps(ref):
    it, _ = ref
    for s in range(500, 512):
        m = synth.synth_label_map(s, 150 + 7 * (s % 5), 170) if s % 3 else synth.synth_noise_label_map(s, 90, 110, block=1 + s % 3)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            want = it.meta_inference(m.astype(np.int64).copy())
            got = mo.meta_inference(m.astype(np.int64).copy())
        assert np.array_equal(want, got), f"seed {s}"
        assert tuple(map(int, it.count_cc(want == 3))) == mo.count_cc(got == 3)


def test_tiling_and_stitch_on_fresh_shapes(ref):
    it, _ = ref
    rng = np.random.default_rng(7)
    for (h, w) in [(333, 777), (512, 512), (700, 301)]:
        img = rng.integers(0, 255, (h, w, 1), dtype=np.uint8)
        _img, patches, pos = it.im2patches_overlap(img)
        pos_o, tiles_o = mo.im2patches_overlap(img)
        assert np.array_equal(np.array(pos), pos_o) and np.array_equal(np.array(patches), tiles_o)
        preds = rng.random((len(pos), 256, 256, 4), dtype=np.float32)
        want = it.patches2im_overlap(list(preds), [list(p) for p in pos])
        assert np.array_equal(want, mo.patches2im_overlap(preds, pos_o))


def test_preprocess_on_fresh_images(ref):
    it, _ = ref
    for s, kw in [(41, {}), (42, {"invert": True}), (43, {"dtype": "u16"}), (44, {"rgb": True, "invert": True})]:
        img = synth.synth_dapi(s, 270, 290, **kw)
        assert np.array_equal(it.meta_preprocess(img.copy()), mo.meta_preprocess(img.copy()))


def test_overlay_helpers_on_fresh_inputs(ref):
    it, _ = ref
    for s in range(60, 66):
        h, w = 140 + 10 * (s % 3), 180
        I = synth.synth_fish(s, h, w, dtype="u16" if s % 2 else "u8")
        seg = synth.synth_label_map(400 + s, h, w).astype(np.int64)
        sens = (85, 30, 180)[s % 3]
        red = it.u16_to_u8(I)[..., 0] > sens
        green = it.u16_to_u8(I)[..., 1] > sens
        nuclei, chrom, ec = seg == 1, seg == 2, seg == 3
        fish, fish2 = green * ~nuclei, red * ~nuclei
        got = mo.overlay_counts(I, seg, sens)
        assert tuple(map(int, it.count_cc(fish * ~chrom))) == got["num_FISH"]
        assert it.count_colocalization(ec, fish) == got["num_ecDNA_FISH"]
        assert it.count_colocalization(fish * ~chrom, fish2 * ~chrom) == got["num_FISH_FISH2"]
        assert it.count_colocalization(ec, fish2 * fish) == got["num_ecDNA_FISH_FISH2"]
        assert it.count_HSR(chrom, fish, 20) == got["num_HSR"]
        assert it.count_HSR(chrom, fish2, 20) == got["num_HSR2"]
