"""runmat_b200 — B200-native (sm_100a) accelerate provider for RunMat's dense-array hot path.

The compute lives in `librm_accel_b200.so` (hand-written CUDA + NVRTC-lowered fused programs) behind the
C ABI in `include/rm_accel.h`. This package is the thin Python host binding used by tests and the bench.
There is no CPU fallback: importing without the built extension raises.
"""
from ._capi import ExtensionMissing, Handle, LIB_PATH  # noqa: F401
from .provider import B200Provider, ImageNormalizeDescriptor, MatmulEpilogue, ProviderError  # noqa: F401

__all__ = ["B200Provider", "ProviderError", "Handle", "MatmulEpilogue", "ImageNormalizeDescriptor", "ExtensionMissing", "LIB_PATH"]
