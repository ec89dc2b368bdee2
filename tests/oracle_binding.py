"""ctypes binding of oracle/librm_oracle.so — the CPU restatement of the reference (checker only)."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "oracle" / "librm_oracle.so"

BINARY_OPS = ["add", "sub", "mul", "div", "pow", "max", "min", "hypot", "atan2", "mod", "rem", "ge", "le", "lt", "gt", "eq", "ne"]
UNARY_OPS = ["sin", "cos", "tan", "asin", "acos", "atan", "sinh", "cosh", "tanh", "asinh", "acosh", "atanh", "exp", "expm1", "log",
             "log2", "log10", "log1p", "sqrt", "abs", "sign", "floor", "ceil", "round", "fix", "neg", "pow2", "heaviside", "single",
             "double", "isnan", "isinf", "isfinite", "nan_to_zero", "not_nan_mask", "erf", "gamma", "gammaln"]
SCALAR_OPS = ["add", "sub", "mul", "div", "rsub", "rdiv", "max", "min", "pow"]

_d = C.POINTER(C.c_double)
_f = C.POINTER(C.c_float)
_u64 = C.POINTER(C.c_uint64)


def _dp(a):
    return a.ctypes.data_as(_d)


def _shape(s):
    return (C.c_uint64 * max(len(s), 1))(*[int(x) for x in s])


def f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1, order="F"))


class Oracle:
    def __init__(self):
        if not LIB.exists():
            raise RuntimeError(f"{LIB} missing: run `make -C oracle`")
        self.lib = C.CDLL(str(LIB))
        L = self.lib
        for name in ("orc_reduce_sum", "orc_reduce_mean", "orc_reduce_prod", "orc_reduce_max", "orc_reduce_min", "orc_mod_scalar",
                     "orc_rem_scalar", "orc_mc_lcg_payoff_sum", "orc_sq_err_sum_f32"):
            getattr(L, name).restype = C.c_double
        for name in ("orc_default_seed", "orc_mix_seed", "orc_advance_state", "orc_generate_uniform", "orc_generate_normal",
                     "orc_stochastic_evolution"):
            getattr(L, name).restype = C.c_uint64
        L.orc_mod_scalar.argtypes = [C.c_double, C.c_double]
        L.orc_rem_scalar.argtypes = [C.c_double, C.c_double]
        L.orc_advance_state.argtypes = [C.c_uint64, C.c_uint64]
        L.orc_mix_seed.argtypes = [C.c_uint64]

    # ---- elementwise -------------------------------------------------------------------------------------
    def elem_binary(self, op: str, a, b):
        a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
        ash, bsh = a.shape, b.shape
        rank = max(len(ash), len(bsh))
        oshape = (C.c_uint64 * rank)()
        r = self.lib.orc_broadcast_shape(_shape(ash), len(ash), _shape(bsh), len(bsh), oshape)
        if r < 0:
            raise ValueError("size mismatch")
        out_shape = tuple(oshape[i] for i in range(r))
        out = np.empty(int(np.prod(out_shape)), dtype=np.float64)
        fa, fb = f64(a), f64(b)
        r2 = self.lib.orc_elem_binary(BINARY_OPS.index(op), _dp(fa), _shape(ash), len(ash), _dp(fb), _shape(bsh), len(bsh), _dp(out), oshape)
        assert r2 == r
        return out.reshape(out_shape, order="F")

    def unary(self, op: str, a):
        a = np.asarray(a, dtype=np.float64)
        fa = f64(a)
        out = np.empty_like(fa)
        self.lib.orc_unary(UNARY_OPS.index(op), _dp(fa), C.c_uint64(fa.size), _dp(out))
        return out.reshape(a.shape, order="F")

    def scalar_op(self, op: str, a, s: float):
        a = np.asarray(a, dtype=np.float64)
        fa = f64(a)
        out = np.empty_like(fa)
        self.lib.orc_scalar_op(SCALAR_OPS.index(op), _dp(fa), C.c_uint64(fa.size), C.c_double(s), _dp(out))
        return out.reshape(a.shape, order="F")

    def sin_mul_add(self, a, b, c: float):
        fa, fb = f64(a), f64(b)
        out = np.empty_like(fa)
        self.lib.orc_sin_mul_add(_dp(fa), _dp(fb), C.c_double(c), C.c_uint64(fa.size), _dp(out))
        return out.reshape(np.asarray(a).shape, order="F")

    def mod(self, a, b): return self.lib.orc_mod_scalar(a, b)
    def rem(self, a, b): return self.lib.orc_rem_scalar(a, b)

    # ---- reductions -------------------------------------------------------------------------------------
    def sum_dims(self, a, dims_zero_based, omit_nan=False):
        a = np.asarray(a, dtype=np.float64)
        mask = (C.c_int * a.ndim)(*[1 if d in dims_zero_based else 0 for d in range(a.ndim)])
        oshape = tuple(1 if d in dims_zero_based else a.shape[d] for d in range(a.ndim))
        out = np.empty(int(np.prod(oshape)), dtype=np.float64)
        fa = f64(a)
        self.lib.orc_sum_dims(_dp(fa), _shape(a.shape), a.ndim, mask, int(omit_nan), _dp(out))
        return out.reshape(oshape, order="F")

    def reduce_sum(self, a): fa = f64(a); return self.lib.orc_reduce_sum(_dp(fa), C.c_uint64(fa.size))
    def reduce_mean(self, a): fa = f64(a); return self.lib.orc_reduce_mean(_dp(fa), C.c_uint64(fa.size))
    def reduce_prod(self, a): fa = f64(a); return self.lib.orc_reduce_prod(_dp(fa), C.c_uint64(fa.size))
    def reduce_max(self, a): fa = f64(a); return self.lib.orc_reduce_max(_dp(fa), C.c_uint64(fa.size))
    def reduce_min(self, a): fa = f64(a); return self.lib.orc_reduce_min(_dp(fa), C.c_uint64(fa.size))

    def reduce_minmax_dim(self, a, dim, is_min):
        a = np.asarray(a, dtype=np.float64)
        rows, cols = a.shape
        n = cols if dim == 0 else rows
        vals, idx = np.empty(n), np.empty(n)
        fa = f64(a)
        self.lib.orc_reduce_minmax_dim(_dp(fa), C.c_uint64(rows), C.c_uint64(cols), dim, int(is_min), _dp(vals), _dp(idx))
        shp = (1, cols) if dim == 0 else (rows, 1)
        return vals.reshape(shp, order="F"), idx.reshape(shp, order="F")

    # ---- linalg -------------------------------------------------------------------------------------------
    def matmul(self, a, b, naive=False):
        a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
        m, k = a.shape
        k2, n = b.shape
        assert k == k2
        out = np.empty(m * n)
        fa, fb = f64(a), f64(b)
        fn = self.lib.orc_matmul_naive if naive else self.lib.orc_matmul
        fn(_dp(fa), C.c_uint64(m), C.c_uint64(k), _dp(fb), C.c_uint64(n), _dp(out))
        return out.reshape((m, n), order="F")

    def matmul_epilogue(self, c, alpha=1.0, beta=0.0, row_scale=None, row_div=False, col_scale=None, col_div=False,
                        clamp_min=None, clamp_max=None, pow_exponent=None, diag=None):
        c = np.asarray(c, dtype=np.float64)
        rows, cols = c.shape
        fc = f64(c).copy()
        rs = f64(row_scale) if row_scale is not None else None
        cs = f64(col_scale) if col_scale is not None else None
        dg = f64(diag).copy() if diag is not None else None
        self.lib.orc_matmul_epilogue(_dp(fc), C.c_uint64(rows), C.c_uint64(cols), C.c_double(alpha), C.c_double(beta),
                                     _dp(rs) if rs is not None else None, int(row_div), _dp(cs) if cs is not None else None, int(col_div),
                                     int(clamp_min is not None), C.c_double(clamp_min or 0.0), int(clamp_max is not None), C.c_double(clamp_max or 0.0),
                                     int(pow_exponent is not None), C.c_double(pow_exponent or 0.0), _dp(dg) if dg is not None else None)
        return fc.reshape((rows, cols), order="F"), dg

    # ---- image -----------------------------------------------------------------------------------------------
    def image_normalize(self, x, eps, gain=None, bias=None, gamma=None, clamp_zero=True, f32=False):
        x = np.asarray(x)
        B, H, W = x.shape
        if f32:
            fx = np.ascontiguousarray(x.astype(np.float32).reshape(-1, order="F"))
            out = np.empty_like(fx)
            self.lib.orc_image_normalize_f32(fx.ctypes.data_as(_f), out.ctypes.data_as(_f), C.c_uint64(B), C.c_uint64(H), C.c_uint64(W), C.c_float(eps),
                                             int(gain is not None), C.c_float(gain or 0), int(bias is not None), C.c_float(bias or 0),
                                             int(gamma is not None), C.c_float(gamma or 0), int(clamp_zero))
        else:
            fx = f64(x)
            out = np.empty_like(fx)
            self.lib.orc_image_normalize(_dp(fx), _dp(out), C.c_uint64(B), C.c_uint64(H), C.c_uint64(W), C.c_double(eps),
                                         int(gain is not None), C.c_double(gain or 0), int(bias is not None), C.c_double(bias or 0),
                                         int(gamma is not None), C.c_double(gamma or 0), int(clamp_zero))
        return out.reshape((B, H, W), order="F")

    def imfilter(self, img, ker, padding="constant", cval=0.0, shape="same", mode="corr", f32=False):
        if f32:
            return self._imfilter_f32(img, ker, padding, cval, shape, mode)
        img, ker = np.asarray(img, dtype=np.float64), np.asarray(ker, dtype=np.float64)
        pad = ["constant", "replicate", "symmetric", "circular"].index(padding)
        shp = ["same", "full", "valid"].index(shape)
        md = ["corr", "conv"].index(mode)
        oshape = (C.c_uint64 * 3)()
        fi, fk = f64(img), f64(ker)
        r = self.lib.orc_imfilter(_dp(fi), _shape(img.shape), img.ndim, _dp(fk), _shape(ker.shape), ker.ndim, pad, C.c_double(cval), shp, md, None, oshape)
        assert r >= 0
        dims = [oshape[0], oshape[1], oshape[2]]
        out = np.empty(int(np.prod(dims)))
        self.lib.orc_imfilter(_dp(fi), _shape(img.shape), img.ndim, _dp(fk), _shape(ker.shape), ker.ndim, pad, C.c_double(cval), shp, md, _dp(out), oshape)
        while len(dims) > max(img.ndim, 2) and dims[-1] == 1:
            dims.pop()
        return out.reshape(dims, order="F")

    def _imfilter_f32(self, img, ker, padding, cval, shape, mode):
        img, ker = np.asarray(img, dtype=np.float32), np.asarray(ker, dtype=np.float32)
        pad = ["constant", "replicate", "symmetric", "circular"].index(padding)
        shp = ["same", "full", "valid"].index(shape)
        md = ["corr", "conv"].index(mode)
        oshape = (C.c_uint64 * 3)()
        fi = np.ascontiguousarray(img.reshape(-1, order="F"))
        fk = np.ascontiguousarray(ker.reshape(-1, order="F"))
        fn = self.lib.orc_imfilter_f32
        r = fn(fi.ctypes.data_as(_f), _shape(img.shape), img.ndim, fk.ctypes.data_as(_f), _shape(ker.shape), ker.ndim, pad, C.c_float(cval), shp, md, None, oshape)
        assert r >= 0
        dims = [oshape[0], oshape[1], oshape[2]]
        out = np.empty(int(np.prod(dims)), dtype=np.float32)
        fn(fi.ctypes.data_as(_f), _shape(img.shape), img.ndim, fk.ctypes.data_as(_f), _shape(ker.shape), ker.ndim, pad, C.c_float(cval), shp, md, out.ctypes.data_as(_f), oshape)
        while len(dims) > max(img.ndim, 2) and dims[-1] == 1:
            dims.pop()
        return out.reshape(dims, order="F")

    # ---- RNG / Monte-Carlo --------------------------------------------------------------------------------------
    def default_seed(self): return self.lib.orc_default_seed()
    def mix_seed(self, s): return self.lib.orc_mix_seed(s)
    def advance_state(self, s, d): return self.lib.orc_advance_state(s, d)

    def generate_uniform(self, state, n):
        out = np.empty(n)
        new = self.lib.orc_generate_uniform(C.c_uint64(state), C.c_uint64(n), _dp(out))
        return out, new

    def generate_normal(self, state, n):
        out = np.empty(n)
        new = self.lib.orc_generate_normal(C.c_uint64(state), C.c_uint64(n), _dp(out))
        return out, new

    def stochastic_evolution_sampled(self, rng_state, s0, length, drift, scale, steps, paths):
        paths = np.ascontiguousarray(paths, dtype=np.uint64)
        out = np.empty(paths.size)
        self.lib.orc_stochastic_evolution_sampled(C.c_uint64(rng_state), C.c_double(s0), C.c_uint64(length), C.c_double(drift), C.c_double(scale),
                                                  C.c_uint32(steps), paths.ctypes.data_as(C.POINTER(C.c_uint64)), C.c_uint64(paths.size), _dp(out))
        return out

    def stochastic_evolution(self, rng_state, data, drift, scale, steps):
        d = f64(data).copy()
        new = self.lib.orc_stochastic_evolution(C.c_uint64(rng_state), _dp(d), C.c_uint64(d.size), C.c_double(drift), C.c_double(scale), C.c_uint32(steps))
        return d.reshape(np.asarray(data).shape, order="F"), new

    def linsolve_triangular(self, t, rhs, lower: bool):
        """Returns (solution, rcond) or raises ValueError('singular') like the host's error."""
        t, rhs = np.asarray(t, dtype=np.float64), np.asarray(rhs, dtype=np.float64)
        n, nrhs = t.shape[0], rhs.shape[1]
        ft_, fr = f64(t), f64(rhs)
        out = np.empty(n * nrhs)
        rc = C.c_double()
        r = self.lib.orc_linsolve_triangular(_dp(ft_), C.c_uint64(n), _dp(fr), C.c_uint64(nrhs), int(lower), _dp(out), C.byref(rc))
        if r != 0:
            raise ValueError("singular")
        return out.reshape((n, nrhs), order="F"), rc.value

    def conv2d(self, sig, ker, mode):
        sig, ker = np.asarray(sig, dtype=np.float64), np.asarray(ker, dtype=np.float64)
        m = ["full", "same", "valid"].index(mode)
        dims = (C.c_uint64 * 2)()
        fs, fk = f64(sig), f64(ker)
        self.lib.orc_conv2d(_dp(fs), C.c_uint64(sig.shape[0]), C.c_uint64(sig.shape[1]), _dp(fk), C.c_uint64(ker.shape[0]), C.c_uint64(ker.shape[1]), m, None, dims)
        out = np.empty(int(dims[0] * dims[1]))
        self.lib.orc_conv2d(_dp(fs), C.c_uint64(sig.shape[0]), C.c_uint64(sig.shape[1]), _dp(fk), C.c_uint64(ker.shape[0]), C.c_uint64(ker.shape[1]), m, _dp(out), dims)
        return out.reshape((dims[0], dims[1]), order="F")

    def power_step_normalize(self, c, eps):
        c = np.asarray(c, dtype=np.float64)
        fc = f64(c).copy()
        self.lib.orc_power_step_normalize(_dp(fc), C.c_uint64(c.shape[0]), C.c_uint64(c.shape[1]), C.c_double(eps))
        return fc.reshape(c.shape, order="F")

    def covariance(self, x, biased=False):
        x = np.asarray(x, dtype=np.float64)
        out = np.empty(x.shape[1] * x.shape[1])
        fx = f64(x)
        self.lib.orc_covariance(_dp(fx), C.c_uint64(x.shape[0]), C.c_uint64(x.shape[1]), int(biased), _dp(out))
        return out.reshape((x.shape[1], x.shape[1]), order="F")

    def linspace(self, start, stop, count):
        out = np.empty(count)
        self.lib.orc_linspace(C.c_double(start), C.c_double(stop), C.c_uint64(count), _dp(out))
        return out.reshape((1, count))

    def transpose(self, a):
        a = np.asarray(a, dtype=np.float64)
        r, c = a.shape
        out = np.empty(r * c)
        fa = f64(a)
        self.lib.orc_transpose(_dp(fa), C.c_uint64(r), C.c_uint64(c), _dp(out))
        return out.reshape((c, r), order="F")

    def mc_lcg_payoff_sum(self, M, T, path0, count, seed=0.0, S0=100.0, mu=0.05, sigma=0.2, dt=1.0 / 252.0, K=100.0):
        return self.lib.orc_mc_lcg_payoff_sum(C.c_uint64(M), C.c_uint32(T), C.c_uint64(path0), C.c_uint64(count), C.c_double(seed), C.c_float(S0),
                                              C.c_float(mu), C.c_float(sigma), C.c_float(dt), C.c_float(K), None)

    def image_lcg_fill(self, B, H, W, seed=0.0, b0=0, bcount=None):
        bcount = B if bcount is None else bcount
        out = np.empty(bcount * H * W, dtype=np.float32)
        self.lib.orc_image_lcg_fill(out.ctypes.data_as(_f), C.c_uint64(B), C.c_uint64(H), C.c_uint64(W), C.c_double(seed), C.c_uint64(b0), C.c_uint64(bcount))
        return out.reshape((bcount, H, W), order="F")
