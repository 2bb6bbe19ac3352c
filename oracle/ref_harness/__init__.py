"""Harness that imports the UNMODIFIED reference modules from /root/reference/src with the
uninstallable libraries (tensorflow, skimage, matplotlib) replaced by the stubs in ./stubs.

TEST INFRASTRUCTURE ONLY, and only usable in the development container: /root/reference does not
exist on the GPU box, so nothing in `-m gpu` tests, smoke() or bench.py calls this.  It is used by
oracle/make_golden.py to freeze golden vectors into tests/golden/ and by the (auto-skipping)
tests/test_reference_live.py.
"""
from __future__ import annotations

import importlib
import os
import sys

REFERENCE_SRC = "/root/reference/src"
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stubs")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_SRC, "image_tools.py"))


def load_reference():
    """Returns (image_tools, utils) -- the reference's own module objects."""
    if not available():
        raise RuntimeError("reference sources not present at " + REFERENCE_SRC)
    for p in (REFERENCE_SRC, _STUBS):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    try:
        import scipy
        if not hasattr(scipy, "misc"):
            import types
            sys.modules.setdefault("scipy.misc", types.ModuleType("scipy.misc"))
            scipy.misc = sys.modules["scipy.misc"]
    except Exception:
        pass
    image_tools = importlib.import_module("image_tools")
    utils = importlib.import_module("utils")
    return image_tools, utils


def install_model_factory(fn) -> None:
    """fn(path) -> object with predict_on_batch; what tf.keras.models.load_model returns."""
    import tensorflow as tf  # the stub
    tf._MODEL_FACTORY["fn"] = fn
