"""fast_racing_b200 — B200-native MINCO / SE(3)-GCOPTER trajectory-optimisation back end.

The product is `libfrx.so` (host C++ + hand-written sm_100a CUDA behind the C ABI of
include/frx.h).  This package only holds the ctypes plumbing (`api`), the in-tree build script
(`build`) and the synthetic corridor generators used by the tests and the benchmark (`synth`).
"""
from . import api, synth  # noqa: F401
from .api import Batch, Problem, load_library  # noqa: F401

__all__ = ["api", "synth", "Batch", "Problem", "load_library"]
