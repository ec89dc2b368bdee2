"""Host-side mirror of `runmat_accelerate_api::AccelProvider` for the hot path, over the C ABI.

Method names, argument meaning and error behaviour follow the trait
(crates/runmat-accelerate-api/src/lib.rs:1386-3152): every failure raises `ProviderError` carrying the
provider's message (the image of `Err(anyhow!(..))`); unimplemented methods raise with
"... not supported by provider" exactly like the trait defaults, so a caller can fall back to host the way
the reference's builtins do. This class is a thin binding: all compute happens in librm_accel_b200.so.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Iterable, Optional, Sequence

import numpy as np

from . import _capi
from ._capi import Handle, lib


class ProviderError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"{_capi.STATUS_NAMES[status] if 0 <= status < 8 else status}: {message}")
        self.status = status
        self.message = message


def _check(status: int) -> None:
    if status != 0:
        raise ProviderError(status, (lib.rm_last_error() or b"").decode("utf-8", "replace"))


def _shape_arr(shape: Sequence[int]):
    arr = (C.c_uint64 * max(len(shape), 1))(*[int(s) for s in shape])
    return arr, len(shape)


@dataclass
class MatmulEpilogue:
    """accelerate-api/src/lib.rs:3498-3550"""

    alpha: float = 1.0
    beta: float = 0.0
    row_scale: Optional[Handle] = None
    col_scale: Optional[Handle] = None
    row_op: str = "multiply"
    col_op: str = "multiply"
    clamp_min: Optional[float] = None
    clamp_max: Optional[float] = None
    pow_exponent: Optional[float] = None
    diag_output: Optional[Handle] = None


@dataclass
class ImageNormalizeDescriptor:
    """accelerate-api/src/lib.rs:3564-3577"""

    batch: int
    height: int
    width: int
    epsilon: float
    gain: Optional[float] = None
    bias: Optional[float] = None
    gamma: Optional[float] = None
    clamp_zero: bool = True


class B200Provider:
    """One CUDA device + stream + buffer table. `precision` is ProviderPrecision ("f64" default)."""

    def __init__(self, cuda_ordinal: int = 0, device_id: int = 0, precision: str = "f64"):
        self._p = C.c_void_p()
        self.precision_name = precision
        _check(lib.rm_provider_create(int(cuda_ordinal), C.c_uint32(device_id), 1 if precision == "f64" else 0, C.byref(self._p)))
        self.np_dtype = np.float64 if precision == "f64" else np.float32

    def close(self) -> None:
        if getattr(self, "_p", None) and self._p.value:
            lib.rm_provider_destroy(self._p)
            self._p = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- identity -------------------------------------------------------------------------------------
    def device_info(self) -> str:
        buf = C.create_string_buffer(512)
        _check(lib.rm_device_info_string(self._p, buf, 512))
        return buf.value.decode()

    # ---- multi-GPU exchange (include/rm_accel.h: rm_comm_*) -------------------------------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = (C.c_uint8 * 128)()
        _check(lib.rm_comm_unique_id(buf, 128))
        return bytes(buf)

    def comm_init(self, unique_id: bytes, rank: int, world: int) -> None:
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id[:128])
        _check(lib.rm_comm_init(self._p, buf, 128, C.c_uint32(rank), C.c_uint32(world)))

    def comm_world_size(self) -> int:
        return int(lib.rm_comm_world_size(self._p))

    def comm_allreduce_sum(self, a: Handle) -> Handle:
        """Sum of `a` over all ranks (new handle; stream-ordered, waited for lazily at first use)."""
        out = Handle()
        _check(lib.rm_comm_allreduce_sum(self._p, C.byref(a), C.byref(out)))
        return out

    def comm_fence(self) -> None:
        _check(lib.rm_comm_fence(self._p))

    def comm_p2p_export(self) -> bytes:
        """Allocates this rank's peer-exchange slot buffer; returns its 64-byte CUDA IPC handle."""
        buf = (C.c_uint8 * 64)()
        _check(lib.rm_comm_p2p_export(self._p, buf, 64))
        return bytes(buf)

    def comm_p2p_connect(self, all_handles: Sequence[bytes], rank: int, world: int) -> None:
        blob = b"".join(h[:64].ljust(64, b"\0") for h in all_handles)
        buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        _check(lib.rm_comm_p2p_connect(self._p, buf, len(blob), C.c_uint32(rank), C.c_uint32(world)))

    def comm_p2p_connected(self) -> bool:
        return bool(lib.rm_comm_p2p_connected(self._p))

    def comm_p2p_error(self) -> int:
        e = C.c_int32()
        _check(lib.rm_comm_p2p_error(self._p, C.byref(e)))
        return int(e.value)

    def fused_reduction_allreduce(self, shader: str, inputs: Sequence[Handle], reduce_len: int, flavor: str = "sum", custom_scale: float = 1.0) -> Handle:
        """Per-rank fused 'all' reduction + sum over the ranks of the peer-memory exchange (publish fused into the kernel tail)."""
        arr = (Handle * len(inputs))(*inputs)
        out = Handle()
        fl = {"sum": 0, "mean": 1, "custom": 2}[flavor]
        _check(lib.rm_fused_reduction_allreduce(self._p, shader.encode(), arr, len(inputs), C.c_uint64(reduce_len), fl, C.c_double(custom_scale), C.byref(out)))
        return out

    def pci_bus_id(self) -> str:
        buf = C.create_string_buffer(32)
        _check(lib.rm_device_pci_bus_id(self._p, buf, 32))
        return buf.value.decode().lower()

    def device_info_struct(self) -> _capi.DeviceInfo:
        info = _capi.DeviceInfo()
        _check(lib.rm_device_info_struct(self._p, C.byref(info)))
        return info

    def device_id(self) -> int:
        return lib.rm_device_id(self._p)

    def precision(self) -> str:
        return "f64" if lib.rm_provider_precision(self._p) == 1 else "f32"

    def synchronize(self) -> None:
        _check(lib.rm_synchronize(self._p))

    def stream(self) -> int:
        s = C.c_void_p()
        _check(lib.rm_get_stream(self._p, C.byref(s)))
        return s.value or 0

    def device_ptr(self, h: Handle) -> tuple[int, int]:
        ptr, n = C.c_void_p(), C.c_uint64()
        _check(lib.rm_device_ptr(self._p, C.byref(h), C.byref(ptr), C.byref(n)))
        return ptr.value or 0, n.value

    def copy_to_device(self, h: Handle, dst_ptr: int, dst_elems: int) -> None:
        _check(lib.rm_copy_to_device(self._p, C.byref(h), C.c_void_p(dst_ptr), C.c_uint64(dst_elems)))

    def warmup(self) -> None:
        _check(lib.rm_warmup(self._p))

    # ---- upload / download / free ---------------------------------------------------------------------
    def upload(self, data, shape: Optional[Sequence[int]] = None) -> Handle:
        """HostTensorView{data:&[f64], shape}: column-major f64 host data."""
        arr = np.asarray(data)
        if shape is None:
            shape = arr.shape if arr.ndim >= 2 else (arr.size, 1) if arr.ndim == 1 else (1, 1)
            arr = np.asfortranarray(arr)
        flat = np.ascontiguousarray(arr.reshape(-1, order="F"))
        h = Handle()
        sa, rank = _shape_arr(shape)
        if flat.dtype == np.float32:
            _check(lib.rm_upload_f32(self._p, flat.ctypes.data_as(C.POINTER(C.c_float)), sa, rank, C.byref(h)))
        else:
            flat = flat.astype(np.float64, copy=False)
            _check(lib.rm_upload(self._p, flat.ctypes.data_as(C.POINTER(C.c_double)), sa, rank, C.byref(h)))
        return h

    def upload_ptr(self, ptr: int, shape: Sequence[int], f32: bool = False) -> Handle:
        h = Handle()
        sa, rank = _shape_arr(shape)
        fn = lib.rm_upload_f32 if f32 else lib.rm_upload
        _check(fn(self._p, C.c_void_p(ptr), sa, rank, C.byref(h)))
        return h

    def download(self, h: Handle, dtype=np.float64) -> np.ndarray:
        """HostTensorOwned: returns a Fortran-ordered array with the handle's shape."""
        n = h.numel
        out = np.empty(n, dtype=dtype)
        if dtype == np.float32:
            _check(lib.rm_download_f32(self._p, C.byref(h), out.ctypes.data_as(C.POINTER(C.c_float)), C.c_uint64(n)))
        else:
            _check(lib.rm_download(self._p, C.byref(h), out.ctypes.data_as(C.POINTER(C.c_double)), C.c_uint64(n)))
        return out.reshape(h.shape, order="F")

    def download_into_ptr(self, h: Handle, ptr: int, n: int, f32: bool = False) -> None:
        fn = lib.rm_download_f32 if f32 else lib.rm_download
        _check(fn(self._p, C.byref(h), C.c_void_p(ptr), C.c_uint64(n)))

    def download_async_into_ptr(self, h: Handle, ptr: int, n: int) -> None:
        _check(lib.rm_download_async(self._p, C.byref(h), C.c_void_p(ptr), C.c_uint64(n)))

    def free(self, h: Handle) -> None:
        _check(lib.rm_free(self._p, C.byref(h)))

    def read_scalar(self, h: Handle, linear_index: int) -> float:
        out = C.c_double()
        _check(lib.rm_read_scalar(self._p, C.byref(h), C.c_uint64(linear_index), C.byref(out)))
        return out.value

    def live_buffers(self) -> int:
        return lib.rm_live_buffers(self._p)

    # ---- constructors / layout ---------------------------------------------------------------------------
    def _ctor(self, fn, shape, *extra) -> Handle:
        h = Handle()
        sa, rank = _shape_arr(shape)
        _check(fn(self._p, sa, rank, *extra, C.byref(h)))
        return h

    def zeros(self, shape): return self._ctor(lib.rm_zeros, shape)
    def ones(self, shape): return self._ctor(lib.rm_ones, shape)
    def fill(self, shape, value: float): return self._ctor(lib.rm_fill, shape, C.c_double(value))
    def eye(self, shape): return self._ctor(lib.rm_eye, shape)
    def zeros_like(self, proto: Handle): return self.zeros(proto.shape)
    def ones_like(self, proto: Handle): return self.ones(proto.shape)
    def random_uniform(self, shape): return self._ctor(lib.rm_random_uniform, shape)
    def random_normal(self, shape): return self._ctor(lib.rm_random_normal, shape)

    def linspace(self, start: float, stop: float, count: int) -> Handle:
        h = Handle()
        _check(lib.rm_linspace(self._p, C.c_double(start), C.c_double(stop), C.c_uint64(count), C.byref(h)))
        return h

    def reshape(self, a: Handle, new_shape) -> Handle:
        h = Handle()
        sa, rank = _shape_arr(new_shape)
        _check(lib.rm_reshape(self._p, C.byref(a), sa, rank, C.byref(h)))
        return h

    def transpose(self, a: Handle) -> Handle:
        h = Handle()
        _check(lib.rm_transpose(self._p, C.byref(a), C.byref(h)))
        return h

    def permute(self, a: Handle, order_zero_based: Sequence[int]) -> Handle:
        h = Handle()
        arr = (C.c_uint32 * len(order_zero_based))(*order_zero_based)
        _check(lib.rm_permute(self._p, C.byref(a), arr, len(order_zero_based), C.byref(h)))
        return h

    def repmat(self, a: Handle, reps: Sequence[int]) -> Handle:
        h = Handle()
        arr = (C.c_uint64 * len(reps))(*reps)
        _check(lib.rm_repmat(self._p, C.byref(a), arr, len(reps), C.byref(h)))
        return h

    def gather_linear(self, source: Handle, indices, output_shape) -> Handle:
        idx = np.ascontiguousarray(indices, dtype=np.uint32)
        h = Handle()
        sa, rank = _shape_arr(output_shape)
        _check(lib.rm_gather_linear(self._p, C.byref(source), idx.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_uint64(idx.size), sa, rank, C.byref(h)))
        return h

    def scatter_linear(self, target: Handle, indices, values: Handle) -> None:
        idx = np.ascontiguousarray(indices, dtype=np.uint32)
        _check(lib.rm_scatter_linear(self._p, C.byref(target), idx.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_uint64(idx.size), C.byref(values)))

    def find(self, a: Handle, limit: Optional[int] = None, direction: str = "first"):
        """ProviderFindResult{linear, rows, cols, values}"""
        outs = [Handle() for _ in range(4)]
        _check(lib.rm_find(self._p, C.byref(a), int(limit is not None), C.c_uint64(limit or 0), int(direction == "last"), *[C.byref(o) for o in outs]))
        return tuple(outs)

    def scatter_column(self, matrix: Handle, col_index: int, values: Handle) -> Handle:
        h = Handle()
        _check(lib.rm_scatter_column(self._p, C.byref(matrix), C.c_uint64(col_index), C.byref(values), C.byref(h)))
        return h

    def scatter_row(self, matrix: Handle, row_index: int, values: Handle) -> Handle:
        h = Handle()
        _check(lib.rm_scatter_row(self._p, C.byref(matrix), C.c_uint64(row_index), C.byref(values), C.byref(h)))
        return h

    def sub2ind(self, dims: Sequence[int], strides: Sequence[int], inputs: Sequence[Handle], scalar_mask: Sequence[bool], length: int, output_shape) -> Handle:
        nd = len(dims)
        arr = (Handle * nd)(*inputs)
        sa, rank = _shape_arr(output_shape)
        h = Handle()
        _check(lib.rm_sub2ind(self._p, (C.c_uint64 * nd)(*dims), (C.c_uint64 * nd)(*strides), nd, arr, (C.c_uint8 * nd)(*[int(b) for b in scalar_mask]),
                              C.c_uint64(length), sa, rank, C.byref(h)))
        return h

    def ind2sub(self, dims: Sequence[int], strides: Sequence[int], indices: Handle, total: int, length: int, output_shape) -> list[Handle]:
        nd = len(dims)
        outs = (Handle * nd)()
        sa, rank = _shape_arr(output_shape)
        _check(lib.rm_ind2sub(self._p, (C.c_uint64 * nd)(*dims), (C.c_uint64 * nd)(*strides), nd, C.byref(indices), C.c_uint64(total), C.c_uint64(length), sa, rank, outs))
        return list(outs)

    # ---- unfused operator surface ----------------------------------------------------------------------------
    def elem_binary(self, op: str, a: Handle, b: Handle) -> Handle:
        h = Handle()
        _check(lib.rm_elem_binary(self._p, _capi.BINARY_OPS.index(op), C.byref(a), C.byref(b), C.byref(h)))
        return h

    def unary(self, op: str, a: Handle) -> Handle:
        h = Handle()
        _check(lib.rm_unary(self._p, _capi.UNARY_OPS.index(op), C.byref(a), C.byref(h)))
        return h

    def scalar_op(self, op: str, a: Handle, scalar: float) -> Handle:
        h = Handle()
        _check(lib.rm_scalar_op_apply(self._p, _capi.SCALAR_OPS.index(op), C.byref(a), C.c_double(scalar), C.byref(h)))
        return h

    def _named2(self, fn, a, b):
        h = Handle()
        _check(fn(self._p, C.byref(a), C.byref(b), C.byref(h)))
        return h

    def _named1(self, fn, a):
        h = Handle()
        _check(fn(self._p, C.byref(a), C.byref(h)))
        return h

    def _nameds(self, fn, a, s):
        h = Handle()
        _check(fn(self._p, C.byref(a), C.c_double(s), C.byref(h)))
        return h

    def elem_add(self, a, b): return self._named2(lib.rm_elem_add, a, b)
    def elem_sub(self, a, b): return self._named2(lib.rm_elem_sub, a, b)
    def elem_mul(self, a, b): return self._named2(lib.rm_elem_mul, a, b)
    def elem_div(self, a, b): return self._named2(lib.rm_elem_div, a, b)
    def elem_pow(self, a, b): return self._named2(lib.rm_elem_pow, a, b)
    def elem_max(self, a, b): return self._named2(lib.rm_elem_max, a, b)
    def elem_min(self, a, b): return self._named2(lib.rm_elem_min, a, b)
    def elem_hypot(self, a, b): return self._named2(lib.rm_elem_hypot, a, b)
    def elem_atan2(self, y, x): return self._named2(lib.rm_elem_atan2, y, x)
    def unary_sin(self, a): return self._named1(lib.rm_unary_sin, a)
    def unary_cos(self, a): return self._named1(lib.rm_unary_cos, a)
    def unary_tan(self, a): return self._named1(lib.rm_unary_tan, a)
    def unary_tanh(self, a): return self._named1(lib.rm_unary_tanh, a)
    def unary_exp(self, a): return self._named1(lib.rm_unary_exp, a)
    def unary_log(self, a): return self._named1(lib.rm_unary_log, a)
    def unary_sqrt(self, a): return self._named1(lib.rm_unary_sqrt, a)
    def unary_abs(self, a): return self._named1(lib.rm_unary_abs, a)
    def unary_floor(self, a): return self._named1(lib.rm_unary_floor, a)
    def unary_round(self, a): return self._named1(lib.rm_unary_round, a)
    def scalar_add(self, a, s): return self._nameds(lib.rm_scalar_add, a, s)
    def scalar_sub(self, a, s): return self._nameds(lib.rm_scalar_sub, a, s)
    def scalar_mul(self, a, s): return self._nameds(lib.rm_scalar_mul, a, s)
    def scalar_div(self, a, s): return self._nameds(lib.rm_scalar_div, a, s)
    def scalar_rsub(self, a, s): return self._nameds(lib.rm_scalar_rsub, a, s)
    def scalar_rdiv(self, a, s): return self._nameds(lib.rm_scalar_rdiv, a, s)
    def scalar_max(self, a, s): return self._nameds(lib.rm_scalar_max, a, s)
    def scalar_min(self, a, s): return self._nameds(lib.rm_scalar_min, a, s)

    # ---- fused ---------------------------------------------------------------------------------------------------
    def fused_elementwise(self, shader: str, inputs: Sequence[Handle], output_shape: Sequence[int], length: int) -> Handle:
        arr = (Handle * len(inputs))(*inputs)
        sa, rank = _shape_arr(output_shape)
        h = Handle()
        _check(lib.rm_fused_elementwise(self._p, shader.encode(), arr, len(inputs), sa, rank, C.c_uint64(length), C.byref(h)))
        return h

    def fused_elementwise_multi(self, shader: str, inputs: Sequence[Handle], output_shape, length: int, num_outputs: int) -> list[Handle]:
        arr = (Handle * len(inputs))(*inputs)
        sa, rank = _shape_arr(output_shape)
        outs = (Handle * num_outputs)()
        _check(lib.rm_fused_elementwise_multi(self._p, shader.encode(), arr, len(inputs), sa, rank, C.c_uint64(length), num_outputs, outs))
        return list(outs)

    def fused_reduction(self, shader: str, inputs: Sequence[Handle], output_shape, reduce_len: int, num_slices: int,
                        workgroup_size: int = 256, flavor: str = "sum", custom_scale: float = 1.0) -> Handle:
        arr = (Handle * len(inputs))(*inputs)
        sa, rank = _shape_arr(output_shape)
        h = Handle()
        fl = {"sum": 0, "mean": 1, "custom": 2}[flavor]
        _check(lib.rm_fused_reduction(self._p, shader.encode(), arr, len(inputs), sa, rank, C.c_uint64(reduce_len), C.c_uint64(num_slices),
                                      C.c_uint32(workgroup_size), fl, C.c_double(custom_scale), C.byref(h)))
        return h

    def fused_cache_counters(self) -> tuple[int, int]:
        hits, misses = C.c_uint64(), C.c_uint64()
        lib.rm_fused_cache_counters(self._p, C.byref(hits), C.byref(misses))
        return hits.value, misses.value

    # ---- reductions -----------------------------------------------------------------------------------------------
    def reduce_sum(self, a): return self._named1(lib.rm_reduce_sum, a)
    def reduce_prod(self, a): return self._named1(lib.rm_reduce_prod, a)
    def reduce_mean(self, a): return self._named1(lib.rm_reduce_mean, a)
    def reduce_max(self, a): return self._named1(lib.rm_reduce_max, a)
    def reduce_min(self, a): return self._named1(lib.rm_reduce_min, a)

    def reduce_sum_dim(self, a, dim: int) -> Handle:
        h = Handle()
        _check(lib.rm_reduce_sum_dim(self._p, C.byref(a), C.c_uint32(dim), C.byref(h)))
        return h

    def reduce_mean_dim(self, a, dim: int) -> Handle:
        h = Handle()
        _check(lib.rm_reduce_mean_dim(self._p, C.byref(a), C.c_uint32(dim), C.byref(h)))
        return h

    def reduce_mean_nd(self, a, dims_zero_based: Iterable[int]) -> Handle:
        d = list(dims_zero_based)
        arr = (C.c_uint32 * len(d))(*d)
        h = Handle()
        _check(lib.rm_reduce_mean_nd(self._p, C.byref(a), arr, len(d), C.byref(h)))
        return h

    def reduce_moments_nd(self, a, dims_zero_based: Iterable[int]) -> tuple[Handle, Handle]:
        d = list(dims_zero_based)
        arr = (C.c_uint32 * len(d))(*d)
        m, e = Handle(), Handle()
        _check(lib.rm_reduce_moments_nd(self._p, C.byref(a), arr, len(d), C.byref(m), C.byref(e)))
        return m, e

    def _minmax_dim(self, fn, a, dim):
        v, i = Handle(), Handle()
        _check(fn(self._p, C.byref(a), C.c_uint32(dim), C.byref(v), C.byref(i)))
        return v, i

    def reduce_max_dim(self, a, dim: int): return self._minmax_dim(lib.rm_reduce_max_dim, a, dim)
    def reduce_min_dim(self, a, dim: int): return self._minmax_dim(lib.rm_reduce_min_dim, a, dim)
    def default_reduction_workgroup_size(self) -> int: return lib.rm_default_reduction_workgroup_size(self._p)
    def two_pass_threshold(self) -> int: return lib.rm_two_pass_threshold(self._p)

    # ---- linalg ---------------------------------------------------------------------------------------------------
    def matmul(self, a, b): return self._named2(lib.rm_matmul, a, b)
    def syrk(self, a): return self._named1(lib.rm_syrk, a)
    def mldivide(self, a, b): return self._named2(lib.rm_mldivide, a, b)

    def mrdivide(self, lhs, rhs): return self._named2(lib.rm_mrdivide, lhs, rhs)

    def linsolve(self, lhs: Handle, rhs: Handle, lower=False, upper=False, rectangular=False, transposed=False, conjugate=False, symmetric=False,
                 posdef=False, need_rcond=False, rcond: Optional[float] = None) -> tuple[Handle, float]:
        """ProviderLinsolveResult{solution, reciprocal_condition} (lib.rs:2422-2429, 694-697)."""
        o = _capi.LinsolveOptions(int(lower), int(upper), int(rectangular), int(transposed), int(conjugate), int(symmetric), int(posdef), int(need_rcond),
                                  int(rcond is not None), float(rcond or 0.0))
        out, rc = Handle(), C.c_double()
        _check(lib.rm_linsolve(self._p, C.byref(lhs), C.byref(rhs), C.byref(o), C.byref(out), C.byref(rc)))
        return out, rc.value

    def conv2d(self, signal: Handle, kernel: Handle, mode: str = "full") -> Handle:
        h = Handle()
        _check(lib.rm_conv2d(self._p, C.byref(signal), C.byref(kernel), ["full", "same", "valid"].index(mode), C.byref(h)))
        return h

    def cat(self, dim_one_based: int, inputs: Sequence[Handle]) -> Handle:
        arr = (Handle * len(inputs))(*inputs)
        h = Handle()
        _check(lib.rm_cat(self._p, C.c_uint32(dim_one_based), arr, len(inputs), C.byref(h)))
        return h

    def matmul_power_step(self, lhs: Handle, rhs: Handle, epsilon: float) -> Handle:
        h = Handle()
        _check(lib.rm_matmul_power_step(self._p, C.byref(lhs), C.byref(rhs), C.c_double(epsilon), C.byref(h)))
        return h

    def covariance(self, matrix: Handle, normalization: str = "unbiased") -> Handle:
        h = Handle()
        _check(lib.rm_covariance(self._p, C.byref(matrix), int(normalization == "biased"), C.byref(h)))
        return h

    def diag_extract(self, matrix: Handle, offset: int = 0) -> Handle:
        h = Handle()
        _check(lib.rm_diag_extract(self._p, C.byref(matrix), C.c_int64(offset), C.byref(h)))
        return h

    def host_sync_count(self) -> int:
        """How many entry-point calls have blocked the host on the device so far (rm_host_sync_count)."""
        return int(lib.rm_host_sync_count(self._p))

    def ozaki_stats(self) -> dict:
        """Device flags of the last tcgen05 product (waits for the stream): test/debug hook."""
        out = (C.c_int32 * 4)()
        _check(lib.rm_debug_ozaki_stats(self._p, out))
        return {"nonfinite": int(out[0]), "pipeline_error": int(out[1]), "fp64_tiles": int(out[2]), "int8_gemms": int(out[3])}

    def kernel_launch_log(self) -> list[dict]:
        """ProviderTelemetry.kernel_launches: the bounded log of recent launches (64 events, newest last)."""
        ev = (_capi.KernelLaunchEvent * 64)()
        n = C.c_uint32()
        _check(lib.rm_kernel_launch_log(self._p, ev, 64, C.byref(n)))
        out = []
        for e in ev[:n.value]:
            out.append({"kernel": e.kernel.decode(), "precision": "f64" if e.precision == 1 else "f32",
                        "shape": {a.key.decode(): int(a.value) for a in e.shape[:e.n_shape]},
                        "tuning": {a.key.decode(): int(a.value) for a in e.tuning[:e.n_tuning]}})
        return out

    def spawn_handle_concurrency(self) -> str:
        """AccelProvider::spawn_handle_concurrency (lib.rs:1400-1402)."""
        return ["ImmutableShare", "CopyOnWrite", "SynchronizedMutation", "Reject"][int(lib.rm_spawn_handle_concurrency_policy(self._p))]

    def set_launch_overlap(self, enabled: bool) -> None:
        """Programmatic dependent launch of the generated fused kernels (default on); off = plain launches (isolated timing)."""
        _check(lib.rm_set_launch_overlap(self._p, 1 if enabled else 0))

    def device_flags(self, n: int = 4) -> list[int]:
        """Device-side pipeline protocol flags (waits for the stream): all zero unless a bounded wait ran out."""
        out = (C.c_int32 * n)()
        _check(lib.rm_debug_device_flags(self._p, out, n))
        return [int(v) for v in out]

    def set_matmul_engine(self, engine: int) -> None:
        _check(lib.rm_set_matmul_engine(self._p, int(engine)))

    def matmul_epilogue(self, a: Handle, b: Handle, ep: MatmulEpilogue) -> Handle:
        c = _capi.MatmulEpilogue()
        c.alpha, c.beta = ep.alpha, ep.beta
        keep = []
        for name in ("row_scale", "col_scale", "diag_output"):
            hv = getattr(ep, name)
            if hv is not None:
                keep.append(hv)
                setattr(c, name, C.pointer(hv))
        c.row_op = 1 if ep.row_op == "divide" else 0
        c.col_op = 1 if ep.col_op == "divide" else 0
        if ep.clamp_min is not None: c.has_clamp_min, c.clamp_min = 1, ep.clamp_min
        if ep.clamp_max is not None: c.has_clamp_max, c.clamp_max = 1, ep.clamp_max
        if ep.pow_exponent is not None: c.has_pow, c.pow_exponent = 1, ep.pow_exponent
        h = Handle()
        _check(lib.rm_matmul_epilogue_apply(self._p, C.byref(a), C.byref(b), C.byref(c), C.byref(h)))
        return h

    # ---- Monte-Carlo / RNG ---------------------------------------------------------------------------------------------
    def set_rng_state(self, state: int) -> None:
        _check(lib.rm_set_rng_state(self._p, C.c_uint64(state)))

    def get_rng_state(self) -> int:
        s = C.c_uint64()
        _check(lib.rm_get_rng_state(self._p, C.byref(s)))
        return s.value

    def stochastic_evolution(self, state: Handle, drift: float, scale: float, steps: int) -> Handle:
        h = Handle()
        _check(lib.rm_stochastic_evolution(self._p, C.byref(state), C.c_double(drift), C.c_double(scale), C.c_uint32(steps), C.byref(h)))
        return h

    def stochastic_evolution_sharded(self, state: Handle, drift: float, scale: float, steps: int, path_offset: int, global_len: int) -> Handle:
        h = Handle()
        _check(lib.rm_stochastic_evolution_sharded(self._p, C.byref(state), C.c_double(drift), C.c_double(scale), C.c_uint32(steps),
                                                   C.c_uint64(path_offset), C.c_uint64(global_len), C.byref(h)))
        return h

    def payoff_partial_sum(self, state: Handle, strike: float) -> Handle:
        h = Handle()
        _check(lib.rm_payoff_partial_sum(self._p, C.byref(state), C.c_double(strike), C.byref(h)))
        return h

    # ---- image ------------------------------------------------------------------------------------------------------------
    def image_normalize(self, a: Handle, d: ImageNormalizeDescriptor) -> Handle:
        c = _capi.ImageNormalizeDesc(d.batch, d.height, d.width, d.epsilon,
                                     int(d.gain is not None), d.gain or 0.0, int(d.bias is not None), d.bias or 0.0,
                                     int(d.gamma is not None), d.gamma or 0.0, int(d.clamp_zero))
        h = Handle()
        _check(lib.rm_image_normalize(self._p, C.byref(a), C.byref(c), C.byref(h)))
        return h

    def imfilter(self, image: Handle, kernel: Handle, padding: str = "constant", constant_value: float = 0.0,
                 shape: str = "same", mode: str = "corr") -> Handle:
        o = _capi.ImfilterOptions(["constant", "replicate", "symmetric", "circular"].index(padding), constant_value,
                                  ["same", "full", "valid"].index(shape), ["corr", "conv"].index(mode))
        h = Handle()
        _check(lib.rm_imfilter(self._p, C.byref(image), C.byref(kernel), C.byref(o), C.byref(h)))
        return h

    # ---- telemetry / measurement -----------------------------------------------------------------------------------------------
    def telemetry_snapshot(self) -> _capi.Telemetry:
        t = _capi.Telemetry()
        _check(lib.rm_telemetry_snapshot(self._p, C.byref(t)))
        return t

    def reset_telemetry(self) -> None:
        _check(lib.rm_reset_telemetry(self._p))

    def timer_begin(self) -> None:
        _check(lib.rm_timer_begin(self._p))

    def timer_end_ms(self) -> float:
        ms = C.c_double()
        _check(lib.rm_timer_end_ms(self._p, C.byref(ms)))
        return ms.value

    def flush_l2(self) -> None:
        _check(lib.rm_flush_l2(self._p))


def pinned_empty(n: int, dtype=np.float64) -> np.ndarray:
    """Page-locked host buffer as a numpy array (for the e2e leg: H2D/D2H from pinned memory)."""
    ptr = C.c_void_p()
    nbytes = int(n) * np.dtype(dtype).itemsize
    _check(lib.rm_pinned_alloc(C.c_size_t(nbytes), C.byref(ptr)))
    buf = (C.c_char * nbytes).from_address(ptr.value)
    arr = np.frombuffer(buf, dtype=dtype, count=n)
    _PINNED.append((ptr, buf))  # freed at process exit
    return arr


_PINNED: list = []
