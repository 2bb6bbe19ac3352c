"""Pipelined metaseg over a list of image files: the loop of the reference's src/metaseg.py:42-57 with decode,
GPU work and file writes overlapped instead of run one after the other.

    reader threads   path -> decoded image in a pinned staging buffer           (tiffio.read_into; src/utils.py:110)
    submit thread    pinned image -> ecseg_segment_image_files_async on one of  (src/utils.py:109-120, src/metaseg.py:46-53)
                     `n_ctx` library contexts / CUDA streams, waits the oldest
    writer threads   write() the three file images the GPU produced             (dapi/<name>, labels/<stem>.png|.npy)

Every slot owns its pinned buffers, so an image's bytes move host -> GPU -> host -> page cache exactly once and no
Python-level pixel loop exists anywhere.  Results come back in input order as (path, n_ec) rows for the CSV."""
from __future__ import annotations

import os
import queue
import threading
import time

import numpy as np
import torch

from . import tiffio
from .engine import Engine


def _pinned(n: int) -> np.ndarray:
    return torch.empty(n, dtype=torch.uint8, pin_memory=True).numpy()


class _Slot:
    def __init__(self, max_h: int, max_w: int, max_bytes_per_px: int):
        png_cap, npy_bytes, tif_bytes = Engine.artifact_sizes(max_h, max_w)
        self.img = _pinned(max_h * max_w * max_bytes_per_px)
        self.tif = _pinned(tif_bytes)
        self.npy = _pinned(npy_bytes)
        self.png = np.empty(png_cap, np.uint8)      # pageable: filled on the host from a pinned staging chunk
        self.view = None


def _get_or_stop(q: "queue.Queue", stop: threading.Event):
    """q.get() that gives up (returns None) once `stop` is set: no thread of the pipeline may block forever on a
    queue whose producer has died (a writer hitting ENOSPC, a GPU call raising, a reader failing to decode)."""
    while True:
        try:
            return q.get(timeout=0.2)
        except queue.Empty:
            if stop.is_set():
                return None


def output_paths(image_path: str):
    """(dapi/<name>, labels/<stem>.png, labels/<stem>.npy) exactly as src/utils.py:122-123 / src/metaseg.py:44-53."""
    d, name = os.path.split(image_path)
    stem = os.path.join(d, 'labels', name[:-4])
    return os.path.join(d, 'dapi', name), stem + '.png', stem + '.npy'


class FilesPipeline:
    def __init__(self, weights: dict, precision: str = "fp16", max_h: int = 2048, max_w: int = 2048, device: int = 0,
                 n_ctx: int = 2, n_slots: int | None = None, n_readers: int = 4, n_writers: int = 6,
                 max_bytes_per_px: int = 8, write_files: bool = True, verbose: bool = False):
        self.max_h, self.max_w = max_h, max_w
        self.engines = [Engine(device, max_h, max_w) for _ in range(n_ctx)]
        for e in self.engines:
            e.load_weights(weights, precision)
        self.streams = [torch.cuda.Stream(device=device) for _ in range(n_ctx)]
        n_slots = n_slots or (n_ctx + n_readers + n_writers)
        self.slots = [_Slot(max_h, max_w, max_bytes_per_px) for _ in range(n_slots)]
        self.n_readers, self.n_writers = n_readers, n_writers
        self.write_files = write_files
        self.verbose = verbose
        self.stats = {}

    def close(self):
        for e in self.engines:
            e.close()
        self.engines = []

    # ------------------------------------------------------------------------------------------
    def run(self, paths):
        """Process `paths`; returns [(path, n_ec)] in input order.  Any worker exception is re-raised."""
        n = len(paths)
        free_q: queue.Queue = queue.Queue()
        for s in self.slots:
            free_q.put(s)
        todo_q: queue.Queue = queue.Queue()
        for i, p in enumerate(paths):
            todo_q.put((i, p))
        decoded_q: queue.Queue = queue.Queue()
        write_q: queue.Queue = queue.Queue()
        results = [None] * n
        errors = []
        stop = threading.Event()
        t_io = {"read": 0.0, "write": 0.0}
        lock = threading.Lock()

        def reader():
            while not stop.is_set():
                try:
                    i, p = todo_q.get_nowait()
                except queue.Empty:
                    return
                slot = _get_or_stop(free_q, stop)
                if slot is None:
                    return
                try:
                    t0 = time.perf_counter()
                    slot.view = tiffio.read_into(p, slot.img)
                    h, w = slot.view.shape[:2]
                    if h > self.max_h or w > self.max_w:
                        raise ValueError(f"{p}: {h}x{w} exceeds the pipeline's {self.max_h}x{self.max_w}")
                    with lock:
                        t_io["read"] += time.perf_counter() - t0
                    decoded_q.put((i, p, slot))
                except BaseException as e:  # noqa: BLE001
                    errors.append(e)
                    stop.set()
                    decoded_q.put(None)
                    return

        def writer():
            while True:
                item = write_q.get()
                if item is None:
                    return
                i, p, slot, n_ec, png_bytes = item
                try:
                    if self.write_files:
                        t0 = time.perf_counter()
                        h, w = slot.view.shape[:2]
                        _, npy_bytes, tif_bytes = Engine.artifact_sizes(h, w)
                        f_tif, f_png, f_npy = output_paths(p)
                        for path, buf, nb in ((f_tif, slot.tif, tif_bytes), (f_png, slot.png, png_bytes), (f_npy, slot.npy, npy_bytes)):
                            with open(path, 'wb', buffering=0) as f:
                                f.write(memoryview(buf[:nb]))
                        with lock:
                            t_io["write"] += time.perf_counter() - t0
                    if self.verbose:
                        print("Processing image: ", p)
                        print("Saving labels: ", p, " to ", output_paths(p)[1][:-4])
                    results[i] = (p, n_ec)
                except BaseException as e:  # noqa: BLE001
                    errors.append(e)
                    stop.set()
                finally:
                    free_q.put(slot)

        readers = [threading.Thread(target=reader, daemon=True) for _ in range(min(self.n_readers, max(n, 1)))]
        writers = [threading.Thread(target=writer, daemon=True) for _ in range(self.n_writers)]
        for t in readers + writers:
            t.start()

        inflight = [None] * len(self.engines)       # per context: (i, path, slot)
        t_start = time.perf_counter()

        def retire(k):
            i, p, slot = inflight[k]
            n_ec, _px, png_bytes = self.engines[k].segment_files_wait()
            inflight[k] = None
            write_q.put((i, p, slot, n_ec, png_bytes))

        try:
            k = 0
            for _ in range(n):
                item = _get_or_stop(decoded_q, stop)     # None: a reader posted its failure, or a worker set `stop`
                if item is None or stop.is_set():
                    break
                i, p, slot = item
                if inflight[k] is not None:
                    retire(k)
                h, w = slot.view.shape[:2]
                _, npy_bytes, tif_bytes = Engine.artifact_sizes(h, w)
                self.engines[k].segment_files_async(slot.view, slot.tif[:tif_bytes], slot.npy[:npy_bytes], slot.png,
                                                    stream=self.streams[k].cuda_stream)
                inflight[k] = (i, p, slot)
                k = (k + 1) % len(self.engines)
            for j in range(len(self.engines)):
                kk = (k + j) % len(self.engines)
                if inflight[kk] is not None:
                    retire(kk)
        except BaseException as e:  # noqa: BLE001
            errors.append(e)
            stop.set()
        finally:
            if errors:      # leave the contexts reusable: drain whatever is still in flight, keep the first error
                for kk, fl in enumerate(inflight):
                    if fl is not None:
                        try:
                            self.engines[kk].segment_files_wait()
                        except BaseException:  # noqa: BLE001
                            pass
                        inflight[kk] = None
            for _ in writers:
                write_q.put(None)
            for t in writers:
                t.join()
            for t in readers:
                t.join(timeout=5)
        if errors:
            raise errors[0]
        wall = time.perf_counter() - t_start
        self.stats = {"images": n, "wall_s": wall, "images_per_s": n / wall if wall > 0 else 0.0,
                      "reader_busy_s": t_io["read"], "writer_busy_s": t_io["write"]}
        return results
