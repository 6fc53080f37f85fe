"""pyribs_b200: the pyribs insertion + CMA-family ask/tell hot path, B200-native.

Importing the package does not need a GPU; constructing an archive or an evolution
strategy does, and raises if the CUDA library under `_C/` is missing (no CPU fallback).
"""
__version__ = "0.1.0"

from pyribs_b200 import archives, emitters, schedulers  # noqa: F401,E402
