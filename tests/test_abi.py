"""CPU-only: the C-ABI library loads and exports every symbol include/ecseg_b200.h declares, the
Python binding table matches the header, and the host-only entry point works without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from ecseg_b200 import _lib, spec

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "ecseg_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ecseg_[a-z0-9_]+)\s*\(", txt)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.isfile(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return _lib.load()


def test_header_and_binding_table_agree():
    assert header_symbols() == sorted(_lib.SIGNATURES)


def test_library_exports_every_declared_symbol(lib):
    for name in header_symbols():
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.ecseg_version()


def test_tile_grid_host_only(lib, golden):
    g = golden("tiling")
    for h, w in g["shapes"]:
        n = ctypes.c_int()
        assert lib.ecseg_tile_grid(int(h), int(w), ctypes.byref(n), None, None, None) == 0
        pos = np.zeros((n.value, 2), np.int32)
        assert lib.ecseg_tile_grid(int(h), int(w), None, None, None, pos.ctypes.data_as(ctypes.c_void_p)) == 0
        assert np.array_equal(pos, g[f"pos_{h}x{w}"]), (h, w)
    assert lib.ecseg_tile_grid(200, 300, None, None, None, None) != 0   # reference cannot tile < 256


def test_no_gpu_means_loud_failure(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ecseg_b200.engine import Engine
    with pytest.raises(RuntimeError):
        Engine(0)
    ctx = ctypes.c_void_p()
    assert lib.ecseg_ctx_create(ctypes.byref(ctx), 0, 256, 256, 1) != 0


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing in the package, the drop-in scripts, the C ABI header or the
    measurement / profiling helpers under tools/ imports or executes it (helpers that need it live in tests/devtools/)."""
    for top in ("ecseg_b200", "src", "include", "tools"):
        for dirpath, _d, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".sh")):
                    txt = open(os.path.join(dirpath, f)).read()
                    assert not re.search(r"^\s*(from|import)\s+oracle|oracle\.[a-z_]+|oracle/|import_module\(.oracle", txt, re.M), \
                        os.path.join(dirpath, f)


def test_flops_and_blob_size():
    assert abs(spec.unet_flops_per_tile() / 1e9 - 97.014) < 0.01
    from ecseg_b200 import weights
    n = sum(9 * l[2] * l[3] + 5 * l[3] + 1 for l in spec.UNET_LAYERS)
    assert spec.n_weight_floats() == n
    assert sum(9 * l[2] * l[3] for l in spec.UNET_LAYERS) == 32148288
