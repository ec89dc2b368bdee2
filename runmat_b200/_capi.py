"""ctypes binding of include/rm_accel.h (the C-ABI drop-in boundary).

The product path has NO CPU fallback: if the CUDA extension is missing this module raises at import,
and `Provider()` raises if no sm_100 device is usable.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

RM_MAX_RANK = 16

_PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("RUNMAT_B200_LIB", _PKG / "librm_accel_b200.so"))


class ExtensionMissing(ImportError):
    pass


if not LIB_PATH.exists():
    raise ExtensionMissing(
        f"{LIB_PATH} not found: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()'). "
        "runmat_b200 has no CPU fallback."
    )

lib = C.CDLL(str(LIB_PATH), mode=C.RTLD_GLOBAL)


class Handle(C.Structure):
    """GpuTensorHandle (accelerate-api/src/lib.rs:260-264)."""

    _fields_ = [("buffer_id", C.c_uint64), ("device_id", C.c_uint32), ("rank", C.c_uint32),
                ("shape_arr", C.c_uint64 * RM_MAX_RANK)]

    @property
    def shape(self) -> tuple[int, ...]:
        return tuple(self.shape_arr[i] for i in range(self.rank))

    @property
    def numel(self) -> int:
        n = 1
        for d in self.shape:
            n *= d
        return n

    def __repr__(self) -> str:
        return f"Handle(buffer_id={self.buffer_id}, device_id={self.device_id}, shape={list(self.shape)})"


class DeviceInfo(C.Structure):
    _fields_ = [("device_id", C.c_uint32), ("name", C.c_char * 128), ("vendor", C.c_char * 32),
                ("backend", C.c_char * 32), ("memory_bytes", C.c_uint64), ("sm_count", C.c_uint32),
                ("cc_major", C.c_uint32), ("cc_minor", C.c_uint32)]


class DispatchStats(C.Structure):
    _fields_ = [("count", C.c_uint64), ("total_wall_time_ns", C.c_uint64)]


class Telemetry(C.Structure):
    _fields_ = [("fused_elementwise", DispatchStats), ("fused_reduction", DispatchStats), ("matmul", DispatchStats),
                ("linsolve", DispatchStats), ("mldivide", DispatchStats), ("mrdivide", DispatchStats),
                ("upload_bytes", C.c_uint64), ("download_bytes", C.c_uint64), ("fusion_cache_hits", C.c_uint64),
                ("fusion_cache_misses", C.c_uint64), ("kernel_launches", C.c_uint64)]


class MatmulEpilogue(C.Structure):
    _fields_ = [("alpha", C.c_double), ("beta", C.c_double), ("row_scale", C.POINTER(Handle)),
                ("col_scale", C.POINTER(Handle)), ("row_op", C.c_int), ("col_op", C.c_int),
                ("has_clamp_min", C.c_int), ("clamp_min", C.c_double), ("has_clamp_max", C.c_int),
                ("clamp_max", C.c_double), ("has_pow", C.c_int), ("pow_exponent", C.c_double),
                ("diag_output", C.POINTER(Handle))]


class ImageNormalizeDesc(C.Structure):
    _fields_ = [("batch", C.c_uint64), ("height", C.c_uint64), ("width", C.c_uint64), ("epsilon", C.c_double),
                ("has_gain", C.c_int), ("gain", C.c_double), ("has_bias", C.c_int), ("bias", C.c_double),
                ("has_gamma", C.c_int), ("gamma", C.c_double), ("clamp_zero", C.c_int)]


class LinsolveOptions(C.Structure):
    """ProviderLinsolveOptions (accelerate-api/src/lib.rs:681-691)."""

    _fields_ = [("lower", C.c_int), ("upper", C.c_int), ("rectangular", C.c_int), ("transposed", C.c_int), ("conjugate", C.c_int),
                ("symmetric", C.c_int), ("posdef", C.c_int), ("need_rcond", C.c_int), ("has_rcond", C.c_int), ("rcond", C.c_double)]


class ImfilterOptions(C.Structure):
    _fields_ = [("padding", C.c_int), ("constant_value", C.c_double), ("shape", C.c_int), ("mode", C.c_int)]


# status codes
RM_OK, RM_ERROR, RM_UNSUPPORTED, RM_OOM, RM_INVALID_HANDLE, RM_INVALID_ARG, RM_NO_DEVICE, RM_COMPILE_ERROR = range(8)
STATUS_NAMES = ["RM_OK", "RM_ERROR", "RM_UNSUPPORTED", "RM_OOM", "RM_INVALID_HANDLE", "RM_INVALID_ARG",
                "RM_NO_DEVICE", "RM_COMPILE_ERROR"]

BINARY_OPS = ["add", "sub", "mul", "div", "pow", "max", "min", "hypot", "atan2", "mod", "rem",
              "ge", "le", "lt", "gt", "eq", "ne"]
UNARY_OPS = ["sin", "cos", "tan", "asin", "acos", "atan", "sinh", "cosh", "tanh", "asinh", "acosh", "atanh",
             "exp", "expm1", "log", "log2", "log10", "log1p", "sqrt", "abs", "sign", "floor", "ceil", "round",
             "fix", "neg", "pow2", "heaviside", "single", "double", "isnan", "isinf", "isfinite",
             "nan_to_zero", "not_nan_mask", "erf", "gamma", "gammaln"]
SCALAR_OPS = ["add", "sub", "mul", "div", "rsub", "rdiv", "max", "min", "pow"]

lib.rm_last_error.restype = C.c_char_p
lib.rm_abi_version.restype = C.c_uint32
lib.rm_device_id.restype = C.c_uint32
lib.rm_live_buffers.restype = C.c_uint64
lib.rm_comm_world_size.restype = C.c_uint32
lib.rm_live_bytes.restype = C.c_uint64
lib.rm_host_sync_count.restype = C.c_uint64
lib.rm_two_pass_threshold.restype = C.c_uint64
lib.rm_default_reduction_workgroup_size.restype = C.c_uint32

class KernelAttr(C.Structure):
    _fields_ = [("key", C.c_char * 16), ("value", C.c_uint64)]


class KernelLaunchEvent(C.Structure):
    """rm_kernel_launch_event (KernelLaunchTelemetry, accelerate-api/src/lib.rs:1369-1375)."""

    _fields_ = [("kernel", C.c_char * 32), ("precision", C.c_int), ("n_shape", C.c_uint32), ("n_tuning", C.c_uint32),
                ("shape", KernelAttr * 4), ("tuning", KernelAttr * 4)]


# Every symbol include/rm_accel.h declares (tests/test_abi.py checks the library exports all of them).
DECLARED_SYMBOLS = None


def declared_symbols() -> list[str]:
    import re

    hdr = (_PKG.parent / "include" / "rm_accel.h").read_text()
    names = re.findall(r"\b(rm_[a-z0-9_]+)\s*\(", hdr)
    seen, out = set(), []
    for n in names:
        if n not in seen:
            seen.add(n)
            out.append(n)
    return out
