"""BASELINE.json config 1 on the CPU: the oracle against tests/golden/example.npz, which
oracle/make_golden.py froze by running the reference's own utils.meta_segment (src/utils.py:109-120)
on input.tif := 255 - example_ecSeg/dapi.jpeg with the fp32 torch-CPU U-Net standing in for Keras."""
import warnings

import numpy as np
import pytest

from oracle import metaseg_oracle as mo


@pytest.fixture(scope="module")
def ex(golden):
    g = golden("example")
    d = {k: g[k] for k in g.files}
    d["notie"] = np.unpackbits(d["notie"])[: 1040 * 1392].reshape(1040, 1392).astype(bool)
    return d


def test_fixture_shape_and_tile_grid(ex):
    assert ex["input"].shape == (1040, 1392) and ex["input"].dtype == np.uint8
    pos, tiles = mo.im2patches_overlap(mo.meta_preprocess(ex["input"])[..., None])
    assert len(pos) == 35 and np.array_equal(pos, ex["pos"])                 # 5 x 7 tiles
    assert not mo.stitch_hole_mask(1040, 1392).any()                         # non-square: image_tools.py:242 never bites
    assert np.array_equal(255 - mo.meta_preprocess(ex["input"]), ex["dapi"])  # utils.py:112


def test_postprocess_and_count_bit_exact_on_reference_raw_map(ex):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = mo.meta_inference(ex["raw"].astype(np.int64).copy())
    assert np.array_equal(out, ex["final"])
    assert mo.count_cc(out == 3) == tuple(int(v) for v in ex["count"])
    assert np.array_equal(np.bincount(out.ravel(), minlength=4), ex["hist_final"])


def test_unet_oracle_reproduces_the_frozen_run(ex):
    """The torch-CPU U-Net on THIS machine against the frozen sparse logits and the frozen label map:
    fp32 summation order may differ between oneDNN builds, so 1e-4 relative / ties excluded."""
    from ecseg_b200 import weights as wmod
    from oracle.unet_oracle import UNetOracle
    net = UNetOracle(wmod.make_weights(0), batch=5)
    pos, tiles = mo.im2patches_overlap(mo.meta_preprocess(ex["input"])[..., None])
    z = net.predict_logits(tiles)
    rel = np.abs(z[:, ::8, ::8, :] - ex["logits_sub8"]).max() / float(ex["logits_absmax"])
    assert rel <= 1e-4, rel
    import torch
    p = torch.softmax(torch.from_numpy(z), -1).numpy()
    raw = mo.quantise_argmax(mo.patches2im_overlap(p, pos))
    agree = float((raw == ex["raw"])[ex["notie"]].mean())
    assert agree >= 0.9999, agree
