"""B200-native streaming reverse-push PPR (drop-in for the hot path of guowentian/dynamicppr).

The product is ``lib/libdppr.so`` (hand-written sm_100a CUDA behind the C ABI of
``include/dppr.h``) and ``bin/pagerank`` (the reference-compatible CLI).  This package is the thin
Python face used by the tests and ``bench.py``: ctypes over the C ABI, nothing else.  There is no
CPU fallback -- :func:`load_library` raises if the CUDA library has not been built.
"""
from .binding import DynamicPPR, DpprError, BatchStats, load_library, library_path  # noqa: F401
from . import graphgen, stream  # noqa: F401

__all__ = ["DynamicPPR", "DpprError", "BatchStats", "load_library", "library_path", "graphgen", "stream"]
