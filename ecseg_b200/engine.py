"""Host-side handle on one GPU's ecseg context: owns the ctypes context, moves numpy/torch buffers
to the device and calls the C ABI.  One Engine per GPU / process (the path shards by image, no
collective: SURVEY.md section 8e)."""
from __future__ import annotations

import ctypes
from ctypes import byref, c_float, c_int, c_int32, c_int64, c_void_p

import numpy as np
import torch

from . import _lib, spec, weights as wmod

PRECISIONS = {"fp32": 0, "bf16": 1, "fp16": 2}
PP_FAITHFUL_MERGE = 1


def tile_grid(h: int, w: int):
    """Tile origins of im2patches_overlap (reference src/image_tools.py:148-186) as int32 [N,2]."""
    lib = _lib.load()
    n, nr, nc = c_int(), c_int(), c_int()
    rc = lib.ecseg_tile_grid(h, w, byref(n), byref(nr), byref(nc), None)
    if rc != 0:
        raise ValueError("images smaller than 256x256 cannot be tiled (reference limitation)")
    pos = np.zeros((n.value, 2), np.int32)
    lib.ecseg_tile_grid(h, w, None, None, None, pos.ctypes.data_as(c_void_p))
    return pos, nr.value, nc.value


def unet_work(h: int, w: int, labels_only: bool = True):
    """(reference FLOPs, executed FLOPs) of one h x w image's U-Net (ecseg_unet_work): the whole-image calls skip the
    blocks of the last four layers that lie in the part of a tile the stitcher never takes."""
    ref, ex = ctypes.c_double(), ctypes.c_double()
    if _lib.load().ecseg_unet_work(h, w, 1 if labels_only else 0, byref(ref), byref(ex)) != 0:
        raise ValueError("images smaller than 256x256 cannot be tiled (reference limitation)")
    return ref.value, ex.value


class Engine:
    def __init__(self, device: int = 0, max_h: int = 2048, max_w: int = 2048, max_tiles: int | None = None):
        if not torch.cuda.is_available():
            raise RuntimeError("ecseg_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        torch.cuda.init()
        if max_tiles is None:
            max_tiles = len(tile_grid(max(max_h, 256), max(max_w, 256))[0])
        self.max_h, self.max_w, self.max_tiles = max_h, max_w, max_tiles
        ctx = c_void_p()
        rc = self.lib.ecseg_ctx_create(byref(ctx), device, max_h, max_w, max_tiles)
        if rc != 0:
            raise _lib.EcsegError(rc, "ecseg_ctx_create failed (out of memory or no such device)")
        self.ctx = ctx
        self.precision = None

    # -- lifetime -------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "ctx", None):
            self.lib.ecseg_ctx_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        _lib.check(self.ctx, rc)

    @staticmethod
    def _stream():
        return c_void_p(torch.cuda.current_stream().cuda_stream)

    def _dev(self, a, dtype=None) -> torch.Tensor:
        if isinstance(a, torch.Tensor):
            t = a.to(self.device)
        else:
            a = np.ascontiguousarray(a)
            if a.dtype == np.uint16:       # torch has limited uint16 support: move raw bytes
                t = torch.from_numpy(a.view(np.int16)).to(self.device)
            else:
                t = torch.from_numpy(a).to(self.device)
        if dtype is not None and t.dtype != dtype:
            t = t.to(dtype)
        return t.contiguous()

    # -- model ----------------------------------------------------------------------------------
    def load_weights(self, w: dict, precision: str = "fp16"):
        blob = wmod.pack_blob(w)
        assert blob.size == spec.n_weight_floats()
        self._chk(self.lib.ecseg_load_weights(self.ctx, blob.ctypes.data_as(c_void_p), blob.size,
                                              PRECISIONS[precision]))
        self.precision = precision

    def debug_set(self, stop_after=-1, tc_cluster=-1, tc_ntile_max=-1):
        """tc_cluster: 0 per-layer table | 1 single CTAs | 2 multicast clusters | 3 CTA pairs; tc_ntile_max: 0 table |
        64 | 128 | 256; -1 leaves a setting unchanged."""
        self._chk(self.lib.ecseg_debug_set(self.ctx, stop_after, tc_cluster, tc_ntile_max))

    def device_error(self) -> int:
        code = c_int()
        self._chk(self.lib.ecseg_device_error(self.ctx, byref(code)))
        return code.value

    def activation_overflow(self) -> int:
        """Index of the first U-Net layer whose 16-bit output held an inf / NaN in the last forward (-1: none).
        Raises EcsegError(ECSEG_E_RANGE) when one did -- the fp16 range guard of the tensor-core modes."""
        layer = c_int()
        self._chk(self.lib.ecseg_activation_overflow(self.ctx, byref(layer)))
        return layer.value

    def launch_count(self) -> int:
        return int(self.lib.ecseg_launch_count(self.ctx))

    # -- stages (device tensors in / out) -------------------------------------------------------
    def preprocess(self, img) -> tuple:
        """meta_preprocess: returns (pre uint8 [H,W], dapi uint8 [H,W]) device tensors."""
        a = img if isinstance(img, torch.Tensor) else np.asarray(img)
        h, w = a.shape[:2]
        ch = 1 if a.ndim == 2 else a.shape[2]
        bps = 2 if (a.dtype in (np.uint16, torch.int16, torch.uint16)) else 1
        d = self._dev(a)
        pre = torch.empty((h, w), dtype=torch.uint8, device=self.device)
        dapi = torch.empty_like(pre)
        self._chk(self.lib.ecseg_preprocess(self.ctx, d.data_ptr(), h, w, ch, bps, pre.data_ptr(), dapi.data_ptr(),
                                            self._stream()))
        return pre, dapi

    def tile(self, pre: torch.Tensor) -> torch.Tensor:
        h, w = pre.shape
        n = len(tile_grid(h, w)[0])
        tiles = torch.empty((n, 256, 256), dtype=torch.uint8, device=self.device)
        self._chk(self.lib.ecseg_tile(self.ctx, pre.data_ptr(), h, w, tiles.data_ptr(), self._stream()))
        return tiles

    def unet_forward(self, tiles, want_logits: bool = False):
        t = self._dev(tiles, torch.uint8).reshape(-1, 256, 256)
        n = t.shape[0]
        probs = torch.empty((n, 256, 256, 4), dtype=torch.float32, device=self.device)
        logits = torch.empty_like(probs) if want_logits else None
        self._chk(self.lib.ecseg_unet_forward(self.ctx, t.data_ptr(), n, probs.data_ptr(),
                                              logits.data_ptr() if want_logits else None, self._stream()))
        return (probs, logits) if want_logits else probs

    def layer_output(self, layer: int, n: int) -> torch.Tensor:
        name, kind, cin, cout, relu, bias, level = spec.UNET_LAYERS[layer]
        hw = 256 >> level
        out = torch.empty((n, hw, hw, cout), dtype=torch.float32, device=self.device)
        self._chk(self.lib.ecseg_debug_layer_output(self.ctx, layer, n, out.data_ptr(), self._stream()))
        return out

    def stitch_argmax(self, probs, h: int, w: int) -> torch.Tensor:
        p = self._dev(probs, torch.float32)
        labels = torch.empty((h, w), dtype=torch.uint8, device=self.device)
        self._chk(self.lib.ecseg_stitch_argmax(self.ctx, p.data_ptr(), h, w, labels.data_ptr(), self._stream()))
        return labels

    def postprocess(self, labels, faithful_merge: bool = False):
        """meta_inference + count_cc(I==3): returns (labels uint8 device tensor, n_ec, ec_px)."""
        lab = self._dev(labels, torch.uint8).clone()
        h, w = lab.shape
        n = torch.zeros(1, dtype=torch.int32, device=self.device)
        px = torch.zeros(1, dtype=torch.int64, device=self.device)
        flags = PP_FAITHFUL_MERGE if faithful_merge else 0
        self._chk(self.lib.ecseg_postprocess(self.ctx, lab.data_ptr(), h, w, flags, n.data_ptr(), px.data_ptr(),
                                             self._stream()))
        return lab, int(n.item()), int(px.item())

    def count_cc(self, mask):
        m = self._dev(np.asarray(mask).astype(np.uint8) if not isinstance(mask, torch.Tensor) else mask, torch.uint8)
        h, w = m.shape
        n = torch.zeros(1, dtype=torch.int32, device=self.device)
        px = torch.zeros(1, dtype=torch.int64, device=self.device)
        self._chk(self.lib.ecseg_count_cc(self.ctx, m.data_ptr(), h, w, n.data_ptr(), px.data_ptr(), self._stream()))
        return int(n.item()), int(px.item())

    def _inplace(self, fn, labels, *args):
        lab = self._dev(labels, torch.uint8).clone()
        h, w = lab.shape
        self._chk(fn(self.ctx, lab.data_ptr(), h, w, *args, self._stream()))
        return lab

    def fill_holes(self, labels, class_id: int):
        return self._inplace(self.lib.ecseg_fill_holes, labels, class_id)

    def size_thresh(self, labels):
        return self._inplace(self.lib.ecseg_size_thresh, labels)

    def merge_comp(self, labels, class_id: int):
        return self._inplace(self.lib.ecseg_merge_comp, labels, class_id)

    def label(self, mask, connectivity: int = 8) -> torch.Tensor:
        m = self._dev(mask, torch.uint8)
        h, w = m.shape
        out = torch.empty((h, w), dtype=torch.int32, device=self.device)
        self._chk(self.lib.ecseg_label(self.ctx, m.data_ptr(), h, w, connectivity, out.data_ptr(), self._stream()))
        return out

    # -- meta_overlay ---------------------------------------------------------------------------
    OVERLAY_OUT = ("n_ecDNA", "px_ecDNA", "n_FISH", "px_FISH", "n_ecDNA_FISH", "n_HSR", "n_FISH2", "px_FISH2",
                   "n_FISH_FISH2", "n_ecDNA_FISH2", "n_ecDNA_FISH_FISH2", "n_HSR2")

    def overlay_counts(self, img, labels, sensitivity: int, want_planes: bool = False):
        """Per-image body of meta_overlay (reference src/meta_overlay.py:59-83) on the GPU.
        Returns the 12 integers of ecseg_overlay_counts as a dict (+ inverted red / green planes)."""
        a = img if isinstance(img, torch.Tensor) else np.asarray(img)
        h, w, ch = a.shape
        bps = 2 if (a.dtype in (np.uint16, torch.int16, torch.uint16)) else 1
        d = self._dev(a)
        lab = self._dev(labels, torch.uint8)
        out = torch.zeros(12, dtype=torch.int64, device=self.device)
        red = torch.empty((h, w), dtype=torch.uint8, device=self.device) if want_planes else None
        green = torch.empty_like(red) if want_planes else None
        self._chk(self.lib.ecseg_overlay_counts(self.ctx, d.data_ptr(), h, w, ch, bps, lab.data_ptr(), int(sensitivity),
                                                red.data_ptr() if want_planes else None,
                                                green.data_ptr() if want_planes else None, out.data_ptr(), self._stream()))
        res = dict(zip(self.OVERLAY_OUT, (int(v) for v in out.cpu().numpy())))
        return (res, red, green) if want_planes else res

    def count_colocalization(self, ob1, ob2) -> int:
        a = self._dev(np.asarray(ob1).astype(np.uint8) if not isinstance(ob1, torch.Tensor) else ob1, torch.uint8)
        b = self._dev(np.asarray(ob2).astype(np.uint8) if not isinstance(ob2, torch.Tensor) else ob2, torch.uint8)
        h, w = a.shape
        out = torch.zeros(1, dtype=torch.int64, device=self.device)
        self._chk(self.lib.ecseg_count_colocalization(self.ctx, a.data_ptr(), b.data_ptr(), h, w, out.data_ptr(), self._stream()))
        return int(out.item())

    def remove_small_objects(self, mask, min_size: int) -> torch.Tensor:
        m = self._dev(np.asarray(mask).astype(np.uint8) if not isinstance(mask, torch.Tensor) else mask, torch.uint8)
        h, w = m.shape
        out = torch.empty_like(m)
        self._chk(self.lib.ecseg_remove_small_objects(self.ctx, m.data_ptr(), h, w, int(min_size), out.data_ptr(), self._stream()))
        return out

    # -- whole image ----------------------------------------------------------------------------
    def segment_device(self, img: torch.Tensor, h: int, w: int, ch: int, bps: int, faithful_merge=False, out=None):
        """Device-resident path: returns (labels, dapi, n_ec tensor, ec_px tensor) without syncing.
        `out` = preallocated (labels, dapi, n, px) device tensors to write into."""
        if out is not None:
            labels, dapi, n, px = out
        else:
            labels = torch.empty((h, w), dtype=torch.uint8, device=self.device)
            dapi = torch.empty_like(labels)
            n = torch.zeros(1, dtype=torch.int32, device=self.device)
            px = torch.zeros(1, dtype=torch.int64, device=self.device)
        self._chk(self.lib.ecseg_segment_image(self.ctx, img.data_ptr(), h, w, ch, bps, dapi.data_ptr(),
                                               labels.data_ptr(), n.data_ptr(), px.data_ptr(),
                                               PP_FAITHFUL_MERGE if faithful_merge else 0, self._stream()))
        return labels, dapi, n, px

    def segment_host(self, img: np.ndarray, labels_out: np.ndarray | None = None, dapi_out: np.ndarray | None = None,
                     faithful_merge=False):
        """Host buffers in / out through ecseg_segment_image_host (the end-to-end call)."""
        img = np.ascontiguousarray(img)
        h, w = img.shape[:2]
        ch = 1 if img.ndim == 2 else img.shape[2]
        bps = img.dtype.itemsize
        if labels_out is None:
            labels_out = np.empty((h, w), np.uint8)
        n, px = c_int32(), c_int64()
        self._chk(self.lib.ecseg_segment_image_host(
            self.ctx, img.ctypes.data_as(c_void_p), h, w, ch, bps,
            dapi_out.ctypes.data_as(c_void_p) if dapi_out is not None else None,
            labels_out.ctypes.data_as(c_void_p), byref(n), byref(px), PP_FAITHFUL_MERGE if faithful_merge else 0))
        return labels_out, n.value, px.value

    def segment_host_async(self, img: np.ndarray, labels_out: np.ndarray, dapi_out: np.ndarray | None = None,
                           faithful_merge=False):
        """Enqueue one image (H2D, whole path, D2H) on the current stream; pair with segment_host_wait().
        `img` / `labels_out` / `dapi_out` must stay alive (ideally pinned) until the wait returns."""
        h, w = img.shape[:2]
        ch = 1 if img.ndim == 2 else img.shape[2]
        self._chk(self.lib.ecseg_segment_image_host_async(
            self.ctx, img.ctypes.data_as(c_void_p), h, w, ch, img.dtype.itemsize,
            dapi_out.ctypes.data_as(c_void_p) if dapi_out is not None else None,
            labels_out.ctypes.data_as(c_void_p), PP_FAITHFUL_MERGE if faithful_merge else 0, self._stream()))

    def segment_host_wait(self):
        n, px = c_int32(), c_int64()
        self._chk(self.lib.ecseg_segment_image_host_wait(self.ctx, byref(n), byref(px)))
        return n.value, px.value

    # -- artefact file images -------------------------------------------------------------------
    @staticmethod
    def artifact_sizes(h: int, w: int):
        """(png worst-case capacity, npy file bytes, tif file bytes) for an h x w image."""
        a, b, c = ctypes.c_size_t(), ctypes.c_size_t(), ctypes.c_size_t()
        if _lib.load().ecseg_artifact_sizes(h, w, byref(a), byref(b), byref(c)) != 0:
            raise ValueError("bad image shape")
        return a.value, b.value, c.value

    def _file_image(self, fn, plane, cap: int) -> np.ndarray:
        d = self._dev(plane, torch.uint8)
        h, w = d.shape
        buf = np.empty(cap, np.uint8)
        n = ctypes.c_size_t()
        self._chk(fn(self.ctx, d.data_ptr(), h, w, buf.ctypes.data_as(c_void_p), cap, byref(n), self._stream()))
        return buf[:n.value]

    def overlay_png(self, labels) -> np.ndarray:
        """labels/<stem>.png of the reference (src/metaseg.py:47-52) as PNG file bytes, encoded on the GPU."""
        h, w = labels.shape
        return self._file_image(self.lib.ecseg_overlay_png, labels, self.artifact_sizes(h, w)[0])

    def labels_npy(self, labels) -> np.ndarray:
        """labels/<stem>.npy of the reference (np.save of the int64 map, src/metaseg.py:53) as file bytes."""
        h, w = labels.shape
        return self._file_image(self.lib.ecseg_labels_npy, labels, self.artifact_sizes(h, w)[1])

    def gray_tiff(self, plane) -> np.ndarray:
        """dapi/<name> of the reference (cv2.imwrite of an 8-bit plane, src/utils.py:122-123) as TIFF file bytes."""
        h, w = plane.shape
        return self._file_image(self.lib.ecseg_gray_tiff, plane, self.artifact_sizes(h, w)[2])

    def segment_files_async(self, img: np.ndarray, tif: np.ndarray | None, npy: np.ndarray | None, png: np.ndarray | None,
                            labels_out: np.ndarray | None = None, faithful_merge=False, stream=None):
        """Enqueue one image: H2D, the whole path, the three file images, D2H.  Buffers are uint8 arrays sized by
        artifact_sizes (tif / npy pinned); pair with segment_files_wait()."""
        h, w = img.shape[:2]
        ch = 1 if img.ndim == 2 else img.shape[2]
        p = lambda a: a.ctypes.data_as(c_void_p) if a is not None else None
        self._chk(self.lib.ecseg_segment_image_files_async(
            self.ctx, p(img), h, w, ch, img.dtype.itemsize, p(tif), p(npy), p(png), png.size if png is not None else 0,
            p(labels_out), PP_FAITHFUL_MERGE if faithful_merge else 0, stream if stream is not None else self._stream()))

    def segment_files_wait(self):
        """-> (n_ec, ec_px, png file bytes) of the image enqueued by segment_files_async."""
        n, px, nb = c_int32(), c_int64(), ctypes.c_size_t()
        self._chk(self.lib.ecseg_segment_image_files_wait(self.ctx, byref(n), byref(px), byref(nb)))
        return n.value, px.value, nb.value

    def last_stage_ms(self):
        ms = (c_float * 4)()
        self._chk(self.lib.ecseg_last_stage_ms(self.ctx, ms))
        return list(ms)
