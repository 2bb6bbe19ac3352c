"""Stub of the `tensorflow` import surface the reference's metaseg path touches (TF 2.8 is not
installable here).  TEST INFRASTRUCTURE ONLY: lets /root/reference/src/utils.py import unmodified.
The model object is injected by the harness; nothing here computes."""
import contextlib
import types


class _Strategy:
    def scope(self):
        return contextlib.nullcontext()


distribute = types.SimpleNamespace(MirroredStrategy=_Strategy)
config = types.SimpleNamespace(list_physical_devices=lambda kind=None: [])
_MODEL_FACTORY = {"fn": None}


def _load_model(path, *a, **k):
    if _MODEL_FACTORY["fn"] is None:
        raise RuntimeError("ref_harness: no model factory installed")
    return _MODEL_FACTORY["fn"](path)


keras = types.SimpleNamespace(models=types.SimpleNamespace(load_model=_load_model))
compat = types.SimpleNamespace(v1=types.SimpleNamespace())
float32 = "float32"
