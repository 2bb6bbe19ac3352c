"""GPU: the artefact encoders and the pipelined driver.

Every file image the GPU hands back must decode (cv2 / numpy / zlib, none of them ours) to exactly what the
reference's writers would have stored for the same label map: plt.imsave with the ListedColormap
(src/metaseg.py:47-52), np.save of the int64 map (src/metaseg.py:53), cv2.imwrite of 255 - I (src/utils.py:112)."""
import ctypes
import io
import os
import subprocess
import zlib

import cv2
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HC = os.path.join(ROOT, "tests", "hostcheck")


@pytest.fixture(scope="module")
def eng():
    from ecseg_b200 import weights as wmod
    from ecseg_b200.engine import Engine
    e = Engine(0, 2048, 2048)
    e.load_weights(wmod.make_weights(0), "fp16")
    yield e
    e.close()


def _decode_png(png):
    img = cv2.imdecode(np.asarray(png), cv2.IMREAD_UNCHANGED)
    assert img is not None and img.ndim == 3 and img.shape[2] == 4
    return img[..., [2, 1, 0, 3]]


def _maps():
    rng = np.random.default_rng(11)
    g = np.load(os.path.join(ROOT, "tests", "golden", "postproc.npz"))
    yield "golden_out0", g["out_0"].astype(np.uint8)
    yield "zeros_257x300", np.zeros((257, 300), np.uint8)
    yield "noise_300x517", rng.integers(0, 4, (300, 517), dtype=np.uint8)       # ragged width, incompressible
    yield "noise_1x1", rng.integers(0, 4, (1, 1), dtype=np.uint8)
    yield "threes_64x2048", np.full((64, 2048), 3, np.uint8)
    yy, xx = np.mgrid[:2048, :2048]
    lab = np.zeros((2048, 2048), np.uint8)
    for _ in range(80):
        cy, cx, r = int(rng.integers(0, 2048)), int(rng.integers(0, 2048)), int(rng.integers(2, 150))
        lab[(yy - cy) ** 2 + (xx - cx) ** 2 < r * r] = rng.integers(1, 4)
    yield "blobs_2048", lab
    yield "noise_2048", rng.integers(0, 4, (2048, 2048), dtype=np.uint8)        # ~19 MB stream: beyond the first D2H chunk


def test_overlay_png_decodes_to_reference_palette_and_equals_host_emulation(eng):
    from oracle import metaseg_oracle as mo
    subprocess.run(["make", "-C", HC], check=True, capture_output=True)
    hc = ctypes.CDLL(os.path.join(HC, "libpngdef_host.so"))
    hc.hostcheck_zlib_stream.restype = ctypes.c_size_t
    hc.hostcheck_zlib_stream.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t]
    hc.hostcheck_zlib_cap.restype = ctypes.c_size_t
    hc.hostcheck_zlib_cap.argtypes = [ctypes.c_int, ctypes.c_int]
    for name, lab in _maps():
        png = eng.overlay_png(lab)
        assert np.array_equal(_decode_png(png), mo.overlay_rgba(lab.astype(np.int64))), name
        h, w = lab.shape
        z = np.asarray(png[41:-16])
        raw = zlib.decompress(z.tobytes())
        assert len(raw) == h * (4 * w + 1), name
        cap = hc.hostcheck_zlib_cap(h, w)
        ref = np.zeros(cap, np.uint8)
        n = hc.hostcheck_zlib_stream(np.ascontiguousarray(lab).ctypes.data, h, w, ref.ctypes.data, cap)
        assert n == z.size and np.array_equal(ref[:n], z), name      # bit-exact with the sequential execution
        if name == "blobs_2048":
            assert png.size < 0.02 * 4 * lab.size                    # ~1 % of the raw RGBA leaves the GPU


def test_labels_npy_is_np_save_of_int64(eng):
    for name, lab in _maps():
        if lab.size < 4:
            continue
        got = eng.labels_npy(lab)
        want = io.BytesIO()
        np.save(want, lab.astype(np.int64))
        assert got.tobytes() == want.getvalue(), name


def test_gray_tiff_decodes(eng):
    rng = np.random.default_rng(2)
    for shape in [(256, 256), (300, 517), (2048, 2048)]:
        plane = rng.integers(0, 256, shape, dtype=np.uint8)
        tif = eng.gray_tiff(plane)
        img = cv2.imdecode(np.asarray(tif), cv2.IMREAD_UNCHANGED)
        assert img is not None and np.array_equal(img, plane), shape


def test_segment_files_equals_segment_host(eng):
    import torch
    from ecseg_b200 import synth
    from ecseg_b200.engine import Engine
    for seed, (h, w), kw in [(31, (300, 340), {}), (32, (462, 300), {"dtype": "u16", "rgb": True}), (33, (2048, 2048), {})]:
        img = synth.synth_dapi(seed, h, w, **kw)
        dapi = np.empty((h, w), np.uint8)
        lab, n_ec, ec_px = eng.segment_host(img, dapi_out=dapi)
        png_cap, npy_bytes, tif_bytes = Engine.artifact_sizes(h, w)
        tif = torch.empty(tif_bytes, dtype=torch.uint8, pin_memory=True).numpy()
        npy = torch.empty(npy_bytes, dtype=torch.uint8, pin_memory=True).numpy()
        png = np.empty(png_cap, np.uint8)
        lab2 = np.empty((h, w), np.uint8)
        eng.segment_files_async(np.ascontiguousarray(img), tif, npy, png, labels_out=lab2)
        n2, px2, png_bytes = eng.segment_files_wait()
        assert (n2, px2) == (n_ec, ec_px) and np.array_equal(lab2, lab)
        back = np.load(io.BytesIO(npy.tobytes()))
        assert back.dtype == np.int64 and np.array_equal(back, lab)
        assert np.array_equal(cv2.imdecode(tif, cv2.IMREAD_UNCHANGED), dapi)
        from ecseg_b200 import spec
        assert np.array_equal(_decode_png(png[:png_bytes]), np.array(spec.PALETTE, np.uint8)[lab])


def test_pipeline_over_mixed_tiffs_equals_per_image_calls(tmp_path, eng):
    from ecseg_b200 import spec, synth, weights as wmod
    from ecseg_b200.pipeline import FilesPipeline, output_paths
    d = tmp_path / "in"
    d.mkdir()
    (d / "dapi").mkdir()
    (d / "labels").mkdir()
    imgs = {}
    for i in range(7):
        kw = [{}, {"invert": True}, {"dtype": "u16"}, {"rgb": True}, {"dtype": "u16", "rgb": True}, {}, {}][i]
        im = synth.synth_dapi(40 + i, 280 + 37 * i, 300 + 29 * (i % 3), **kw)
        p = str(d / f"img{i}.tif")
        bgr = im[..., ::-1] if im.ndim == 3 else im
        flags = [cv2.IMWRITE_TIFF_COMPRESSION, 1] if i % 2 == 0 else []      # odd ones LZW: general-decoder fallback
        assert cv2.imwrite(p, np.ascontiguousarray(bgr), flags)
        imgs[p] = im
    pipe = FilesPipeline(wmod.make_weights(0), "fp16", 512, 512, n_ctx=2, n_readers=3, n_writers=3)
    try:
        rows = pipe.run(list(imgs))
    finally:
        pipe.close()
    assert [r[0] for r in rows] == list(imgs)
    for (p, n_ec), im in zip(rows, imgs.values()):
        dapi = np.empty(im.shape[:2], np.uint8)
        lab, n, _ = eng.segment_host(im, dapi_out=dapi)
        f_tif, f_png, f_npy = output_paths(p)
        assert n_ec == n, p
        back = np.load(f_npy)
        assert back.dtype == np.int64 and np.array_equal(back, lab), p
        assert np.array_equal(cv2.imread(f_tif, cv2.IMREAD_UNCHANGED), dapi), p
        png = cv2.imread(f_png, cv2.IMREAD_UNCHANGED)
        assert np.array_equal(png[..., [2, 1, 0, 3]], np.array(spec.PALETTE, np.uint8)[lab]), p
