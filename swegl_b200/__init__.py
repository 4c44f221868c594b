"""swegl_b200 — B200-native implementation of swegl's per-frame rendering hot path.

The product is the CUDA library libswegl_b200.so (swegl_b200/csrc, C ABI in include/swegl_b200.h)
and the C++ drop-in adapter in swegl_b200/host/.  This Python package is the thin harness used by
tests/ and bench.py: ctypes bindings, the flattened scene container and the host-side camera math.
"""
from . import _abi
from .scene import Scene, Viewport, Camera
from .renderer import Renderer

__all__ = ["Scene", "Viewport", "Camera", "Renderer", "_abi"]
