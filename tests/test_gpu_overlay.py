"""GPU parity for meta_overlay: ecseg_overlay_counts, ecseg_count_colocalization,
ecseg_remove_small_objects vs the golden vectors generated from the reference and vs the CPU oracle; the
`python src/meta_overlay.py` surface.  Integer work: exact."""
import os
import subprocess
import sys

import cv2
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORDER = ("n_ecDNA", "px_ecDNA", "n_FISH", "px_FISH", "n_ecDNA_FISH", "n_HSR", "n_FISH2", "px_FISH2", "n_FISH_FISH2",
         "n_ecDNA_FISH2", "n_ecDNA_FISH_FISH2", "n_HSR2")


@pytest.fixture(scope="module")
def eng():
    from ecseg_b200.engine import Engine
    e = Engine(0, 2048, 2048, max_tiles=0)
    yield e
    e.close()


def _oracle_flat(I, seg, sens):
    from oracle import metaseg_oracle as mo
    r = mo.overlay_counts(I, seg, sens)
    return [r["num_ecDNA"][0], r["num_ecDNA"][1], r["num_FISH"][0], r["num_FISH"][1], r["num_ecDNA_FISH"], r["num_HSR"],
            r["num_FISH2"][0], r["num_FISH2"][1], r["num_FISH_FISH2"], r["num_ecDNA_FISH2"], r["num_ecDNA_FISH_FISH2"],
            r["num_HSR2"]]


def test_overlay_golden(eng, golden):
    g = golden("overlay")
    for i in range(int(g["n_cases"])):
        res = eng.overlay_counts(g[f"img_{i}"], g[f"seg_{i}"], int(g[f"sens_{i}"]))
        assert [res[k] for k in ORDER] == [int(v) for v in g[f"out_{i}"]], i


def test_overlay_full_size_and_planes_vs_oracle(eng):
    """BASELINE config 5 shape: 2048x2048 RGB (u8 and u16) with green / red FISH spots."""
    from ecseg_b200 import synth
    from oracle import metaseg_oracle as mo
    for seed, dtype, sens in [(3, "u8", 85), (4, "u16", 120)]:
        I = synth.synth_fish(seed, 2048, 2048, dtype=dtype)
        seg = synth.synth_label_map(700 + seed, 2048, 2048)
        res, red, green = eng.overlay_counts(I, seg, sens, want_planes=True)
        assert [res[k] for k in ORDER] == _oracle_flat(I, seg, sens), seed
        _r, _g, red_inv, green_inv = mo.split_FISH_channels(I, sens)
        assert np.array_equal(red.cpu().numpy(), red_inv) and np.array_equal(green.cpu().numpy(), green_inv)


def test_function_level_helpers(eng):
    from ecseg_b200 import synth
    from oracle import metaseg_oracle as mo
    rng = np.random.default_rng(5)
    for seed, shape in [(0, (97, 131)), (1, (300, 520)), (2, (33, 1000))]:
        a = synth.synth_noise_label_map(seed, *shape, block=2) == 2
        b = rng.random(shape) < 0.02
        assert eng.count_colocalization(a, b) == mo.count_colocalization(a, b)
        for min_size in (1, 5, 20):
            got = eng.remove_small_objects(a, min_size).cpu().numpy().astype(bool)
            assert np.array_equal(got, mo.remove_small_objects(a, min_size)), (seed, min_size)
        big = mo.remove_small_objects(b | a, 20)
        assert eng.count_colocalization(a, big) == mo.count_HSR(a, b | a, 20)
    full = np.ones((40, 50), bool)
    assert eng.count_colocalization(full, full) == 0          # np.unique(regs)[1:] drops the only component
    assert eng.count_colocalization(np.zeros((40, 50), bool), full) == 0


def test_meta_overlay_cli(tmp_path):
    """`python src/meta_overlay.py` after metaseg-style labels exist: CSV header/rows, red/green planes, exit codes."""
    from ecseg_b200 import synth
    from oracle import metaseg_oracle as mo
    data = tmp_path / "data"
    (data / "labels").mkdir(parents=True)
    (tmp_path / "config.yaml").write_text(f"meta_overlay:\n  inpath: {data}\n  color_sensitivity: 85\n")
    env = dict(os.environ, PYTHONPATH=ROOT)
    run = lambda: subprocess.run([sys.executable, os.path.join(ROOT, "src", "meta_overlay.py")], cwd=str(tmp_path), env=env,
                                 capture_output=True, text=True, timeout=600)
    r = run()
    assert r.returncode == 2 and "`dapi` folder is missing" in r.stdout        # meta_overlay.py:30-33
    (data / "dapi").mkdir()
    I = synth.synth_fish(9, 300, 320)
    seg = synth.synth_label_map(909, 300, 320)
    cv2.imwrite(str(data / "x.tif"), I[..., ::-1])
    np.save(data / "labels" / "x.npy", seg.astype(np.int64))
    r = run()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = (data / "fish_quantification.csv").read_text().strip().splitlines()
    assert lines[0] == ("image_name,# of ecDNA (DAPI),# of ecDNA (green),# of ecDNA (red),# of ecDNA (DAPI and green),"
                        "# of ecDNA (DAPI and red),# of ecDNA (red and green),# of ecDNA (DAPI and red and green),"
                        "# of HSR (red),# of HSR (green)")
    o = mo.overlay_counts(I, seg, 85)
    cc = lambda t: f'"({t[0]}, {t[1] if t[1] else 0.0})"'
    want = ",".join(["x.tif", cc(o["num_ecDNA"]), cc(o["num_FISH"]), cc(o["num_FISH2"]), str(o["num_ecDNA_FISH"]),
                     str(o["num_ecDNA_FISH2"]), str(o["num_FISH_FISH2"]), str(o["num_ecDNA_FISH_FISH2"]),
                     str(o["num_HSR2"]), str(o["num_HSR"])])
    assert lines[1] == want
    red = cv2.imread(str(data / "red" / "x.tif.png"), cv2.IMREAD_UNCHANGED)
    assert np.array_equal(red, 255 - I[..., 0])


def test_config5_chain_metaseg_then_meta_overlay(tmp_path):
    """BASELINE config 5 as the user runs it: `make metaseg` then `make meta_overlay` over the same folder of synthetic
    RGB DAPI + green / red FISH images (DAPI in the blue channel, src/image_tools.py:88-89).  metaseg's
    labels/<stem>.npy feeds meta_overlay through read_seg (src/utils.py:125-132); all nine count columns of
    fish_quantification.csv must equal the oracle's overlay_counts on the label map metaseg wrote."""
    from ecseg_b200 import synth
    from oracle import metaseg_oracle as mo
    data = tmp_path / "data"
    data.mkdir()
    imgs = {"f0.tif": synth.synth_fish(21, 462, 470), "f1.tif": synth.synth_fish(22, 300, 330, dtype="u16")}
    for name, I in imgs.items():
        cv2.imwrite(str(data / name), I[..., ::-1])
    (tmp_path / "config.yaml").write_text(f"metaseg:\n  inpath: {data}\nmeta_overlay:\n  inpath: {data}\n  color_sensitivity: 85\n")
    env = dict(os.environ, PYTHONPATH=ROOT, ECSEG_ALLOW_RANDOM_WEIGHTS="1")
    for script in ("metaseg.py", "meta_overlay.py"):
        r = subprocess.run([sys.executable, os.path.join(ROOT, "src", script)], cwd=str(tmp_path), env=env, capture_output=True,
                           text=True, timeout=900)
        assert r.returncode == 0, (script, r.stderr[-2000:])
    lines = (data / "fish_quantification.csv").read_text().strip().splitlines()
    rows = {l.split(",", 1)[0]: l for l in lines[1:]}
    assert set(rows) == set(imgs)
    cc = lambda t: f'"({t[0]}, {t[1] if t[1] else 0.0})"'
    ec_csv = dict(l.rsplit(",", 1) for l in (data / "ec_quantification.csv").read_text().strip().splitlines()[1:])
    for name, I in imgs.items():
        seg = np.load(data / "labels" / (name[:-4] + ".npy"))
        assert seg.dtype == np.int64 and seg.shape == I.shape[:2]
        o = mo.overlay_counts(I, seg.astype(np.uint8), 85)
        want = ",".join([name, cc(o["num_ecDNA"]), cc(o["num_FISH"]), cc(o["num_FISH2"]), str(o["num_ecDNA_FISH"]),
                         str(o["num_ecDNA_FISH2"]), str(o["num_FISH_FISH2"]), str(o["num_ecDNA_FISH_FISH2"]),
                         str(o["num_HSR2"]), str(o["num_HSR"])])
        assert rows[name] == want, name
        assert int(ec_csv[name]) == o["num_ecDNA"][0], name        # the two CSVs agree on the ecDNA count (metaseg.py:46)
