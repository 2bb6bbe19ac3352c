"""Host logic of ecseg_b200.pipeline.FilesPipeline with the GPU contexts replaced by fakes (CPU only): results come
back in input order, and a failing writer / GPU call / reader surfaces as an exception instead of a hang
(the submit loop used to block forever in decoded_q.get() once a writer had died)."""
import os
import threading
import time

import cv2
import numpy as np
import pytest

from ecseg_b200 import pipeline as pl
from ecseg_b200.engine import Engine


class FakeEngine:
    """segment_files_async / segment_files_wait of Engine: 'count' = first pixel of the image."""

    def __init__(self, fail_on=None):
        self.pending, self.fail_on, self.waits = None, fail_on, 0

    def segment_files_async(self, img, tif, npy, png, labels_out=None, faithful_merge=False, stream=None):
        if self.fail_on is not None and int(img.flat[0]) == self.fail_on:
            raise RuntimeError("fake CUDA failure")
        tif[:4] = 1; npy[:4] = 2; png[:4] = 3
        self.pending = int(img.flat[0])

    def segment_files_wait(self):
        self.waits += 1
        n, self.pending = self.pending, None
        return n, 0, 4

    def close(self):
        pass


class FakeSlot:
    def __init__(self, h, w):
        png_cap, npy_bytes, tif_bytes = Engine.artifact_sizes(h, w)
        self.img = np.empty(h * w, np.uint8)
        self.tif, self.npy, self.png = np.empty(tif_bytes, np.uint8), np.empty(npy_bytes, np.uint8), np.empty(png_cap, np.uint8)
        self.view = None


class _FakeStream:
    cuda_stream = 0


def make_pipe(n_ctx=2, n_slots=5, fail_on=None, h=256, w=256):
    p = pl.FilesPipeline.__new__(pl.FilesPipeline)
    p.max_h, p.max_w = h, w
    p.engines = [FakeEngine(fail_on) for _ in range(n_ctx)]
    p.streams = [_FakeStream() for _ in range(n_ctx)]
    p.slots = [FakeSlot(h, w) for _ in range(n_slots)]
    p.n_readers, p.n_writers, p.write_files, p.verbose, p.stats = 2, 2, True, False, {}
    return p


def write_inputs(d, n, make_dirs=True):
    if make_dirs:
        os.mkdir(os.path.join(d, "dapi")); os.mkdir(os.path.join(d, "labels"))
    paths = []
    for i in range(n):
        p = os.path.join(d, f"im{i:02d}.tif")
        cv2.imwrite(p, np.full((256, 256), i, np.uint8), [cv2.IMWRITE_TIFF_COMPRESSION, 1])
        paths.append(p)
    return paths


def run_with_deadline(fn, seconds=20):
    box = {}

    def target():
        try:
            box["result"] = fn()
        except BaseException as e:  # noqa: BLE001
            box["error"] = e

    t = threading.Thread(target=target, daemon=True)
    t0 = time.time()
    t.start()
    t.join(seconds)
    assert not t.is_alive(), f"pipeline still blocked after {seconds} s"
    return box, time.time() - t0


def test_results_in_input_order_and_files_written(tmp_path):
    paths = write_inputs(str(tmp_path), 9)
    pipe = make_pipe()
    box, _ = run_with_deadline(lambda: pipe.run(paths))
    assert "error" not in box, box.get("error")
    assert box["result"] == [(p, i) for i, p in enumerate(paths)]
    for p in paths:
        for f in pl.output_paths(p):
            assert os.path.getsize(f) > 0
    assert pipe.stats["images"] == 9


def test_writer_failure_raises_instead_of_hanging(tmp_path):
    paths = write_inputs(str(tmp_path), 12, make_dirs=False)     # no dapi/ labels/: every open() fails
    pipe = make_pipe()
    box, took = run_with_deadline(lambda: pipe.run(paths))
    assert isinstance(box.get("error"), OSError), box
    assert took < 15


def test_gpu_call_failure_raises_and_drains_inflight(tmp_path):
    paths = write_inputs(str(tmp_path), 12)
    pipe = make_pipe(fail_on=5)
    box, _ = run_with_deadline(lambda: pipe.run(paths))
    assert isinstance(box.get("error"), RuntimeError) and "fake CUDA failure" in str(box["error"])
    assert all(e.pending is None for e in pipe.engines)          # nothing left in flight: contexts reusable


def test_reader_failure_raises(tmp_path):
    paths = write_inputs(str(tmp_path), 6)
    paths.insert(3, os.path.join(str(tmp_path), "missing.tif"))
    pipe = make_pipe()
    box, _ = run_with_deadline(lambda: pipe.run(paths))
    assert isinstance(box.get("error"), FileNotFoundError)
