"""Drop-in for the metaseg part of the reference's src/image_tools.py: same function names,
argument meaning and return types, computed by libecseg_b200 on the GPU.

    reference function (src/image_tools.py)        C ABI entry point
    meta_preprocess            :86-96              ecseg_preprocess
    im2patches_overlap         :148-186            ecseg_tile / ecseg_tile_grid
    patches2im_overlap + img_as_ubyte + argmax     ecseg_stitch_argmax   (stitch_argmax below)
    meta_inference             :15-84              ecseg_postprocess
    count_cc                   :114-119            ecseg_count_cc
    count_colocalization       :126-134            ecseg_count_colocalization
    count_HSR                  :103-112            ecseg_remove_small_objects + ecseg_count_colocalization
    split_FISH_channels        :136-146            (thresholds on the host here; fused in ecseg_overlay_counts)
"""
from __future__ import annotations

import numpy as np

from . import spec
from .engine import Engine, tile_grid

NUM_CLASSES = spec.NUM_CLASSES
EC_SIZE_THRESHOLD = spec.EC_SIZE_THRESHOLD

_engines: dict = {}      # CUDA device index -> Engine


def default_engine(max_h: int = 2048, max_w: int = 2048, device: int | None = None) -> Engine:
    """Per-process engine on `device` (default: torch's CURRENT CUDA device, so that a rank that called
    torch.cuda.set_device(local_rank) stays on its own GPU), created on first use and regrown when an image is
    taller OR wider than what it was sized for."""
    import torch
    if device is None:
        if not torch.cuda.is_available():
            raise RuntimeError("ecseg_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        device = torch.cuda.current_device()
    max_h, max_w = max(max_h, 256), max(max_w, 256)
    eng = _engines.get(device)
    if eng is None or eng.max_h < max_h or eng.max_w < max_w:
        keep = None
        if eng is not None:
            keep = (eng._weights, eng.precision) if getattr(eng, "_weights", None) else None
            max_h, max_w = max(max_h, eng.max_h), max(max_w, eng.max_w)
            eng.close()
        eng = Engine(device, max_h, max_w)
        if keep:
            eng.load_weights(*keep)
            eng._weights = keep[0]
        _engines[device] = eng
    return eng


def close_engines() -> None:
    """Release every cached engine (its ~7 GB of activation workspace) -- used before handing the GPU to the
    overlapped pipeline's own contexts."""
    for eng in _engines.values():
        eng.close()
    _engines.clear()


def meta_preprocess(img: np.ndarray) -> np.ndarray:
    """uint8/uint16, gray or RGB(A) -> uint8 [H,W] with dark background (reference :86-96)."""
    eng = default_engine(*img.shape[:2])
    pre, _ = eng.preprocess(img)
    return pre.cpu().numpy()


def im2patches_overlap(img: np.ndarray, overlap_value: int = 25, scw: int = 256):
    """[img, list of 256x256(x1) tiles, list of [row, col] origins] (reference :148-186)."""
    if overlap_value != spec.OVERLAP or scw != spec.TILE:
        raise ValueError("the metaseg path is fixed at overlap_value=25, scw=256")
    a = img[..., 0] if img.ndim == 3 else img
    eng = default_engine(*a.shape)
    tiles = eng.tile(eng._dev(a)).cpu().numpy()
    pos, _, _ = tile_grid(*a.shape)
    if img.ndim == 3:
        tiles = tiles[..., None]
    return [img, list(tiles), [list(map(int, p)) for p in pos]]


def stitch_argmax(preds, L_pos=None, shape=None) -> np.ndarray:
    """patches2im_overlap (:188-252) + img_as_ubyte + np.argmax (src/utils.py:116-118), fused:
    float32 [N,256,256,4] -> int64 [H,W].  Raises ValueError like img_as_ubyte when a value is
    outside [-1, 1]."""
    if shape is None:
        pos = np.asarray(L_pos)
        shape = (int(pos[:, 0].max()) + spec.TILE, int(pos[:, 1].max()) + spec.TILE)
    eng = default_engine(*shape)
    return eng.stitch_argmax(np.asarray(preds, np.float32), *shape).cpu().numpy().astype(np.int64)


def meta_inference(img: np.ndarray) -> np.ndarray:
    """Post-process a 4-class label map IN PLACE and return it (reference :15-84)."""
    eng = default_engine(*img.shape)
    out, _, _ = eng.postprocess(img.astype(np.uint8))
    img[...] = out.cpu().numpy().astype(img.dtype)
    return img


def count_cc(I: np.ndarray):
    """(number of 8-connected components, pixel total) of a boolean mask (reference :114-119)."""
    eng = default_engine(*I.shape)
    return eng.count_cc(np.asarray(I) != 0)


def count_colocalization(ob1: np.ndarray, ob2: np.ndarray) -> int:
    """Number of 8-connected components of ob1 holding at least one pixel of ob2 (reference :126-134)."""
    eng = default_engine(*ob1.shape)
    return eng.count_colocalization(np.asarray(ob1) != 0, np.asarray(ob2) != 0)


def count_HSR(chrom: np.ndarray, fish: np.ndarray, HSR_SIZE_THRESHOLD: int) -> int:
    """Chromosome components touched by FISH signal left after remove_small_objects (reference :103-112)."""
    eng = default_engine(*chrom.shape)
    big = eng.remove_small_objects(np.asarray(fish) != 0, HSR_SIZE_THRESHOLD)
    return eng.count_colocalization(eng._dev((np.asarray(chrom) != 0).astype(np.uint8)), big)
