"""BASELINE.json config 1 on the GPU: input.tif := 255 - example_ecSeg/dapi.jpeg (1040x1392, 35 tiles -- the only
image the reference ships) with seed-0 weights.  tests/golden/example.npz holds what the reference's own
utils.meta_segment (src/utils.py:109-120) returned for it with the fp32 torch-CPU U-Net as the model object
(oracle/make_golden.py gen_example), so nothing here reads /root/reference.

Bars (BASELINE.json north_star): fp32 logits within 1e-3 relative, labels >= 99.9 % excluding quantised top-2 ties,
post-processing and ecDNA count bit-exact given the same label map."""
import os
import subprocess
import sys
import warnings

import cv2
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
H, W = 1040, 1392


@pytest.fixture(scope="module")
def ex(golden):
    g = golden("example")
    d = {k: g[k] for k in g.files}
    d["notie"] = np.unpackbits(d["notie"])[: H * W].reshape(H, W).astype(bool)
    return d


@pytest.fixture(scope="module")
def eng():
    from ecseg_b200.engine import Engine
    e = Engine(0, H, W)
    yield e
    e.close()


def test_front_end_bit_exact(ex, eng):
    from ecseg_b200.engine import tile_grid
    pre, dapi = eng.preprocess(ex["input"])
    assert np.array_equal(dapi.cpu().numpy(), ex["dapi"])                      # what utils.py:112 wrote to dapi/
    pos, nr, nc = tile_grid(H, W)
    assert (nr, nc) == (5, 7) and np.array_equal(pos, ex["pos"])               # image_tools.py:148-186


def test_postprocess_and_count_bit_exact_on_reference_raw_map(ex, eng):
    lab, n, px = eng.postprocess(ex["raw"])
    assert np.array_equal(lab.cpu().numpy(), ex["final"])                      # image_tools.py:15-84
    assert (n, px) == tuple(int(v) for v in ex["count"])                       # metaseg.py:46
    lab2, n2, px2 = eng.postprocess(ex["raw"], faithful_merge=True)            # with the two merge_comp passes run
    assert np.array_equal(lab2.cpu().numpy(), ex["final"]) and (n2, px2) == (n, px)


def test_fp32_logits_and_labels_on_all_35_tiles(ex, eng):
    import torch
    from ecseg_b200 import weights as wmod
    from oracle import metaseg_oracle as mo
    from oracle.unet_oracle import UNetOracle
    w = wmod.make_weights(0)
    torch.set_num_threads(os.cpu_count() or 1)
    pos, tiles = mo.im2patches_overlap(mo.meta_preprocess(ex["input"])[..., None])
    z_ref = UNetOracle(w, batch=5).predict_logits(tiles)
    # the oracle on this box is the oracle that made the fixture
    assert np.abs(z_ref[:, ::8, ::8, :] - ex["logits_sub8"]).max() / float(ex["logits_absmax"]) <= 1e-4
    eng.load_weights(w, "fp32")
    pre, _ = eng.preprocess(ex["input"])
    probs, logits = eng.unet_forward(eng.tile(pre), want_logits=True)
    rel = float(np.abs(logits.cpu().numpy() - z_ref).max() / np.abs(z_ref).max())
    assert rel <= 1e-3, rel
    raw = eng.stitch_argmax(probs, H, W).cpu().numpy()
    agree = float((raw == ex["raw"])[ex["notie"]].mean())
    print(f"config 1 fp32: logits rel {rel:.2e}, labels {agree * 100:.4f}%")
    assert agree >= 0.999, agree


@pytest.mark.parametrize("prec,bar", [("fp16", 0.999), ("bf16", 0.995)])
def test_tensor_core_labels_vs_reference_run(ex, eng, prec, bar):
    from ecseg_b200 import weights as wmod
    eng.load_weights(wmod.make_weights(0), prec)
    labels, n_ec, ec_px = eng.segment_host(ex["input"])
    pre, _ = eng.preprocess(ex["input"])
    raw = eng.stitch_argmax(eng.unet_forward(eng.tile(pre)), H, W).cpu().numpy()
    assert eng.activation_overflow() == -1
    agree = float((raw == ex["raw"])[ex["notie"]].mean())
    print(f"config 1 {prec}: labels {agree * 100:.4f}% (ties excluded), n_ec {n_ec} (reference run: {int(ex['count'][0])})")
    assert agree >= bar, agree
    # exact count GIVEN the same label map (the bar of north_star): post-process the GPU's own raw map on the CPU oracle
    from oracle import metaseg_oracle as mo
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = mo.meta_inference(raw.astype(np.int64).copy())
    assert np.array_equal(labels, want) and (n_ec, ec_px) == mo.count_cc(want == 3)


def test_metaseg_cli_on_the_example(ex, tmp_path):
    """`python src/metaseg.py` (== make metaseg) over a folder holding input.tif writes the four artefacts of
    src/metaseg.py:47-57 / src/utils.py:112."""
    data = tmp_path / "example_ecSeg"
    data.mkdir()
    cv2.imwrite(str(data / "input.tif"), ex["input"])
    (tmp_path / "config.yaml").write_text(f"metaseg:\n  inpath: {data}\n")
    env = dict(os.environ, PYTHONPATH=ROOT, ECSEG_ALLOW_RANDOM_WEIGHTS="1", ECSEG_PRECISION="fp32")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "src", "metaseg.py")], cwd=str(tmp_path), env=env,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    assert np.array_equal(cv2.imread(str(data / "dapi" / "input.tif"), cv2.IMREAD_UNCHANGED), ex["dapi"])
    npy = np.load(data / "labels" / "input.npy")
    assert npy.dtype == np.int64 and npy.shape == (H, W)
    # fp32 parity mode reproduces the reference run's final map except where a near-tie flipped the raw label
    assert float((npy == ex["final"]).mean()) >= 0.999
    png = cv2.imread(str(data / "labels" / "input.png"), cv2.IMREAD_UNCHANGED)
    assert png.shape == (H, W, 4)
    rows = (data / "ec_quantification.csv").read_text().strip().splitlines()
    assert rows[0] == "image name,# of ec" and rows[1].startswith("input.tif,")
    from oracle import metaseg_oracle as mo
    assert int(rows[1].split(",")[1]) == mo.count_cc(npy == 3)[0]
