"""ctypes binding of libecseg_b200.so (include/ecseg_b200.h).

There is NO CPU fallback: if the library is missing or no CUDA device is present, every call
fails loudly.  PyTorch is used only to own device memory and streams.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# ECSEG_B200_LIB: another build of the same library (A/B runs of two kernel versions on one GPU box)
LIB_PATH = os.environ.get("ECSEG_B200_LIB") or os.path.join(_HERE, "libecseg_b200.so")

# name -> (restype, argtypes).  Every symbol declared in include/ecseg_b200.h is listed here and
# tests/test_abi.py checks the two stay in sync.
SIGNATURES = {
    "ecseg_ctx_create": (c_int, [POINTER(c_void_p), c_int, c_int, c_int, c_int]),
    "ecseg_ctx_destroy": (None, [c_void_p]),
    "ecseg_last_error": (c_char_p, [c_void_p]),
    "ecseg_version": (c_char_p, []),
    "ecseg_load_weights": (c_int, [c_void_p, c_void_p, c_size_t, c_int]),
    "ecseg_tile_grid": (c_int, [c_int, c_int, POINTER(c_int), POINTER(c_int), POINTER(c_int), c_void_p]),
    "ecseg_preprocess": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "ecseg_tile": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "ecseg_unet_forward": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "ecseg_stitch_argmax": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "ecseg_postprocess": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "ecseg_count_cc": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "ecseg_fill_holes": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "ecseg_size_thresh": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "ecseg_merge_comp": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "ecseg_label": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "ecseg_overlay_counts": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p,
                                     c_void_p, c_void_p]),
    "ecseg_count_colocalization": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "ecseg_remove_small_objects": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "ecseg_segment_image": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_int, c_void_p]),
    "ecseg_segment_image_host": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                         POINTER(c_int32), POINTER(c_int64), c_int]),
    "ecseg_segment_image_host_async": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                               c_int, c_void_p]),
    "ecseg_segment_image_host_wait": (c_int, [c_void_p, POINTER(c_int32), POINTER(c_int64)]),
    "ecseg_artifact_sizes": (c_int, [c_int, c_int, POINTER(c_size_t), POINTER(c_size_t), POINTER(c_size_t)]),
    "ecseg_overlay_png": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_size_t, POINTER(c_size_t), c_void_p]),
    "ecseg_labels_npy": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_size_t, POINTER(c_size_t), c_void_p]),
    "ecseg_gray_tiff": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_size_t, POINTER(c_size_t), c_void_p]),
    "ecseg_segment_image_files_async": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                                c_void_p, c_size_t, c_void_p, c_int, c_void_p]),
    "ecseg_segment_image_files_wait": (c_int, [c_void_p, POINTER(c_int32), POINTER(c_int64), POINTER(c_size_t)]),
    "ecseg_tiff_read": (c_int, [c_char_p, c_void_p, c_size_t, POINTER(c_int), POINTER(c_int), POINTER(c_int),
                                POINTER(c_int)]),
    "ecseg_png_wrap": (c_size_t, [c_void_p, c_size_t, c_int, c_int]),
    "ecseg_npy_header": (c_size_t, [c_void_p, c_size_t, c_int, c_int]),
    "ecseg_tiff_header": (c_size_t, [c_void_p, c_int, c_int]),
    "ecseg_crc32": (ctypes.c_uint32, [ctypes.c_uint32, c_void_p, c_size_t]),
    "ecseg_debug_layer_output": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "ecseg_debug_set": (c_int, [c_void_p, c_int, c_int, c_int]),
    "ecseg_debug_progress": (c_int, [c_void_p, c_void_p]),
    "ecseg_debug_trace": (c_int, [c_void_p, c_void_p, c_int]),
    "ecseg_device_error": (c_int, [c_void_p, POINTER(c_int)]),
    "ecseg_activation_overflow": (c_int, [c_void_p, POINTER(c_int)]),
    "ecseg_debug_owned_blocks": (c_int, [c_int, c_int, c_int, c_void_p, POINTER(c_int), POINTER(c_int)]),
    "ecseg_unet_work": (c_int, [c_int, c_int, c_int, POINTER(ctypes.c_double), POINTER(ctypes.c_double)]),
    "ecseg_launch_count": (c_int64, [c_void_p]),
    "ecseg_last_stage_ms": (c_int, [c_void_p, POINTER(c_float)]),
}

_lib = None


class EcsegError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libecseg_b200 error {code}: {msg}")
        self.code = code


def load() -> ctypes.CDLL:
    """Load the shared library (building it is __graft_entry__.build() / make -C ecseg_b200/csrc)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C ecseg_b200/csrc`.  ecseg_b200 has no CPU fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)   # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(ctx, rc: int) -> None:
    if rc != 0:
        msg = load().ecseg_last_error(ctx)
        msg = msg.decode() if msg else ""
        if rc == -4:
            raise ValueError(msg)   # img_as_ubyte's ValueError in the reference (src/utils.py:117)
        raise EcsegError(rc, msg)
