"""ratatosk_b200 — B200-native (sm_100a) implementation of Ratatosk's long-read correction hot path.

The product is the C-ABI shared library `librtk_b200.so` (include/rtk.h, sources in csrc/);
this package is the thin host-side binding used by the tests and bench.py.
"""
from .api import (Context, Graph, RtkError, RtkOpt, default_opt, load_library, pack_reads)  # noqa: F401

__all__ = ["Context", "Graph", "RtkError", "RtkOpt", "default_opt", "load_library", "pack_reads"]
