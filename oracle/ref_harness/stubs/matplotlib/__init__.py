"""Stub of the two matplotlib entry points on the reference's metaseg path (src/metaseg.py:47-52).
TEST INFRASTRUCTURE ONLY."""
from . import pyplot, colors  # noqa: F401
