from . import rank  # noqa: F401
