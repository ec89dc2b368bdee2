"""Parity of the CUDA provider (through the C ABI) against the CPU oracle. Run with -m gpu on a B200.

Tolerances (stated per BASELINE.json north_star):
  * IEEE-exact ops (+ - * / sqrt, rounding, sign, abs, mod/rem, comparisons, all layout/indexing): BIT-EXACT.
    The kernels are compiled with -fmad=false, so no a*b+c contraction changes a rounding.
  * libm-backed ops (sin, exp, pow, ...): |got-want| <= 1e-10*|want| + 1e-13. CUDA's and glibc's implementations are
    each within ~1-2 ulp of the true value but not bit-identical; the absolute floor covers cancellation to ~0.
  * reductions / matmul: order of accumulation differs from the sequential host loop: 1e-10 relative to the
    magnitude bound of the sum (sum |terms|).
"""
import math
import os

import numpy as np
import pytest

from kat_util import KATS, arr, assert_same, num
from runmat_b200 import ImageNormalizeDescriptor, MatmulEpilogue, ProviderError
import fusion_text as ft

pytestmark = pytest.mark.gpu

EXACT_BIN = ["add", "sub", "mul", "div", "max", "min", "mod", "rem", "ge", "le", "lt", "gt", "eq", "ne"]
LIBM_BIN = ["pow", "hypot", "atan2"]
EXACT_UN = ["sqrt", "abs", "sign", "floor", "ceil", "round", "fix", "neg", "heaviside", "single", "double", "isnan", "isinf", "isfinite",
            "nan_to_zero", "not_nan_mask"]
LIBM_UN = ["erf", "gamma", "gammaln", "sin", "cos", "tan", "asin", "acos", "atan", "sinh", "cosh", "tanh", "asinh", "acosh", "atanh", "exp", "expm1", "log", "log2",
           "log10", "log1p", "pow2"]


def close(got, want, rtol=1e-10, atol=1e-13):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (got.shape, want.shape)
    ng, nw = np.isnan(got), np.isnan(want)
    assert np.array_equal(ng, nw), "NaN pattern differs"
    inf = np.isinf(want)
    assert np.array_equal(got[inf], want[inf])
    m = ~(nw | inf)
    err = np.abs(got[m] - want[m])
    bound = rtol * np.abs(want[m]) + atol
    assert np.all(err <= bound), f"max excess {np.max(err - bound)}, worst rel {np.max(err / np.maximum(np.abs(want[m]), 1e-300))}"


def special_values():
    return np.array([0.0, -0.0, 1.0, -1.0, 0.5, -0.5, 2.5, -2.5, 1e-300, 1e300, np.inf, -np.inf, np.nan, 3.0, -7.25, 1e-8])


# ---------------------------------------------------------------------------------------------------------------
# a1/a2: handles, upload / download / free
# ---------------------------------------------------------------------------------------------------------------
def test_upload_download_roundtrip_and_free(prov):
    rng = np.random.default_rng(0)
    for shape in [(1, 1), (3, 1), (1, 5), (7, 9), (4, 3, 2), (0, 3), (1025, 3)]:
        x = rng.uniform(-1, 1, shape)
        h = prov.upload(x)
        assert tuple(h.shape) == shape and h.device_id == 7
        assert np.array_equal(prov.download(h), x)
        prov.free(h)
    n0 = prov.live_buffers()
    h = prov.upload(np.ones((2, 2)))
    assert prov.live_buffers() == n0 + 1
    prov.free(h)
    prov.free(h)  # double free of an unknown id is a no-op (simple_provider.rs:2761)
    assert prov.live_buffers() == n0
    with pytest.raises(ProviderError):
        prov.download(h)
    h.device_id = 99
    with pytest.raises(ProviderError, match="belongs to device"):
        prov.free(h)


def test_pipelined_upload_compute_download(prov, orc):
    """Uploads ride a dedicated H2D stream and rm_download_async skips the final wait: a chunked, software-pipelined pass over
    host buffers must equal the one-shot result (this is the e2e leg of bench.py)."""
    from runmat_b200.provider import pinned_empty

    n, chunks = 1 << 20, 8
    rng = np.random.default_rng(31)
    hostA, hostB, hostC = pinned_empty(n), pinned_empty(n), pinned_empty(n)
    hostA[:] = rng.uniform(0, 6, n)
    hostB[:] = rng.uniform(-1, 1, n)
    hostC[:] = 0.0
    one = prov.upload(np.array([[1.0]]))
    c = n // chunks
    pending = None
    for i in range(chunks + 1):
        nxt = None
        if i < chunks:
            nxt = (i, prov.upload_ptr(hostA.ctypes.data + i * c * 8, (c, 1)), prov.upload_ptr(hostB.ctypes.data + i * c * 8, (c, 1)))
        if pending is not None:
            j, pa, pb, pc = pending
            prov.download_async_into_ptr(pc, hostC.ctypes.data + j * c * 8, c)
            for h in (pa, pb, pc):
                prov.free(h)
        if nxt is not None:
            pending = (nxt[0], nxt[1], nxt[2], prov.fused_elementwise(ft.sin_mul_add_wgsl(), [nxt[1], nxt[2], one], (c, 1), c))
        else:
            pending = None
    prov.synchronize()
    ha, hb = prov.upload(np.array(hostA).reshape(-1, 1)), prov.upload(np.array(hostB).reshape(-1, 1))
    whole = prov.download(prov.fused_elementwise(ft.sin_mul_add_wgsl(), [ha, hb, one], (n, 1), n))[:, 0]
    assert np.array_equal(np.array(hostC), whole)
    close(np.array(hostC), orc.sin_mul_add(np.array(hostA), np.array(hostB), 1.0))


def test_concurrent_host_threads(prov, orc):
    """The provider is Send + Sync (accelerate-api lib.rs:1386): several host threads (think GC finalizer + VM) issue
    uploads, ops, downloads and frees concurrently."""
    import threading

    errors = []

    def worker(seed):
        try:
            rng = np.random.default_rng(seed)
            for _ in range(20):
                a, b = rng.uniform(-1, 1, (257, 13)), rng.uniform(-1, 1, (257, 13))
                ha, hb = prov.upload(a), prov.upload(b)
                hc = prov.elem_mul(ha, hb)
                hs = prov.reduce_sum_dim(hc, 0)
                got_c, got_s = prov.download(hc), prov.download(hs)
                assert np.array_equal(got_c, a * b)
                assert np.allclose(got_s, (a * b).sum(axis=0, keepdims=True), rtol=1e-13, atol=1e-13)
                for h in (ha, hb, hc, hs):
                    prov.free(h)
        except Exception as e:  # noqa
            errors.append(e)

    n0 = prov.live_buffers()
    ts = [threading.Thread(target=worker, args=(s,)) for s in range(6)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors
    assert prov.live_buffers() == n0


def test_precision_and_device_info(prov, prov32):
    assert prov.precision() == "f64" and prov32.precision() == "f32"
    info = prov.device_info_struct()
    assert info.cc_major >= 10 and info.sm_count > 0 and b"cuda" in info.backend
    assert "runmat-b200" in prov.device_info()
    x = np.array([[1.0 + 1e-12, 2.0]])
    h = prov32.upload(x)
    assert np.array_equal(prov32.download(h), x.astype(np.float32).astype(np.float64))  # narrows on upload (io.rs:87)
    prov32.free(h)


def test_reshape_is_metadata_only_and_read_scalar(prov):
    x = np.arange(12.0).reshape((3, 4), order="F")
    h = prov.upload(x)
    r = prov.reshape(h, (2, 6))
    assert r.buffer_id == h.buffer_id
    assert np.array_equal(prov.download(r), x.reshape((2, 6), order="F"))
    assert prov.read_scalar(h, 5) == 5.0
    with pytest.raises(ProviderError):
        prov.reshape(h, (5, 5))
    prov.free(h)


# ---------------------------------------------------------------------------------------------------------------
# a5: unfused operator surface
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("op", EXACT_BIN + LIBM_BIN)
def test_elem_binary(prov, orc, op):
    rng = np.random.default_rng(1)
    cases = [(rng.uniform(-3, 3, (37, 29)), rng.uniform(-3, 3, (37, 29))),       # same shape (flat path, ragged tail)
             (rng.uniform(-3, 3, (33, 1)), rng.uniform(-3, 3, (1, 17))),         # row x column broadcast
             (rng.uniform(-3, 3, (8, 5, 3)), rng.uniform(0.5, 3, (1, 5, 1))),    # N-D broadcast
             (rng.uniform(0.1, 3, (64, 64)), np.array([[2.0]]))]                 # tensor x scalar tensor
    sv = special_values()
    cases.append((sv.reshape(-1, 1) * np.ones((1, sv.size)), np.ones((sv.size, 1)) * sv.reshape(1, -1)))
    for a, b in cases:
        if op == "pow":
            a = np.abs(a)
        ha, hb = prov.upload(a), prov.upload(b)
        hc = prov.elem_binary(op, ha, hb)
        got, want = prov.download(hc), orc.elem_binary(op, a, b)
        if op in EXACT_BIN:
            assert_same(got, want)
        else:
            close(got, want)
        for h in (ha, hb, hc):
            prov.free(h)


def test_elem_binary_kats_and_errors(prov):
    for k in KATS["elem_binary"]:
        ha, hb = prov.upload(arr(k["a"])), prov.upload(arr(k["b"]))
        hc = prov.elem_binary(k["op"], ha, hb)
        assert_same(prov.download(hc), arr(k["out"]))
        for h in (ha, hb, hc):
            prov.free(h)
    ha, hb = prov.upload(np.ones((2, 3))), prov.upload(np.ones((3, 2)))
    with pytest.raises(ProviderError, match="size mismatch"):
        prov.elem_add(ha, hb)
    he = prov.upload(np.ones((0, 3)))
    hz = prov.elem_add(he, prov.upload(np.ones((1, 3))))
    assert tuple(hz.shape) == (0, 3)


@pytest.mark.parametrize("name", ["mod", "rem"])
def test_mod_rem_kats(prov, name):
    a = np.array([[num(r[0]) for r in KATS[name]]])
    b = np.array([[num(r[1]) for r in KATS[name]]])
    want = np.array([[num(r[2]) for r in KATS[name]]])
    ha, hb = prov.upload(a), prov.upload(b)
    hc = prov.elem_binary(name, ha, hb)
    assert_same(prov.download(hc), want)


@pytest.mark.parametrize("op", EXACT_UN + LIBM_UN)
def test_unary(prov, orc, op):
    rng = np.random.default_rng(2)
    x = rng.uniform(-4, 4, (61, 47))
    if op in ("log", "log2", "log10", "sqrt"):
        x = np.abs(x) + 1e-3
    if op in ("asin", "acos", "atanh"):
        x = x / 4.001
    if op == "acosh":
        x = np.abs(x) + 1.0
    if op == "log1p":
        x = np.abs(x)
    if op in ("gamma", "gammaln"):
        x = np.abs(x) + 0.1
    x.flat[: special_values().size] = special_values() if op in EXACT_UN else x.flat[: special_values().size]
    h = prov.upload(x)
    hr = prov.unary(op, h)
    got, want = prov.download(hr), orc.unary(op, x)
    if op in EXACT_UN:
        assert_same(got, want)
    else:
        close(got, want)
    prov.free(h)
    prov.free(hr)


@pytest.mark.parametrize("op", ["sin", "cos"])
def test_trig_ranges_and_special_values(prov, orc, op):
    """The generated kernels' own f64 sin / cos (fusion_lower.cpp prelude, "rm_trig") against the oracle (glibc, what f64::sin /
    f64::cos call: sin.rs:265-270): the fast range up to 105615, arguments next to the zeros of both functions (relative
    accuracy there needs the three-part pi), the out-of-line library route for huge arguments, the fdlibm answers for tiny ones,
    NaN for Inf / NaN, and the sign of sin(-0)."""
    rng = np.random.default_rng(33)
    k = np.arange(-2000, 2001, dtype=np.float64)
    near = np.concatenate([np.nextafter(k * (np.pi / 2), np.inf), k * (np.pi / 2), np.nextafter(k * (np.pi / 2), -np.inf)])
    x = np.concatenate([rng.uniform(-4 * np.pi, 4 * np.pi, 20000), rng.uniform(-105615, 105615, 20000), near,
                        rng.uniform(-1, 1, 2000) * 1e-9, 10.0 ** rng.uniform(5, 300, 2000) * rng.choice([-1.0, 1.0], 2000),
                        np.array([0.0, -0.0, 5e-324, -5e-324, 1e-300, 2.0 ** -27, -(2.0 ** -27), 105614.999, 105615.0, 105615.001, 1e22, -1e22,
                                  np.inf, -np.inf, np.nan, 1.7976931348623157e308])])
    h = prov.upload(x.reshape(-1, 1))
    got, want = prov.download(prov.unary(op, h)).ravel(), orc.unary(op, x.reshape(-1, 1)).ravel()
    close(got, want, rtol=1e-10, atol=0.0)   # purely relative: also next to the zeros
    zero = want == 0.0
    assert np.array_equal(np.signbit(got[zero]), np.signbit(want[zero]))
    if op == "cos":
        assert np.all(got[x == 0.0] == 1.0)
    prov.free(h)


@pytest.mark.parametrize("op", ["add", "sub", "mul", "div", "rsub", "rdiv", "max", "min"])
def test_scalar_ops_exact(prov, orc, op):
    x = np.random.default_rng(3).uniform(-5, 5, (129, 3))
    h = prov.upload(x)
    hr = prov.scalar_op(op, h, 1.7)
    assert_same(prov.download(hr), orc.scalar_op(op, x, 1.7))
    prov.free(h)
    prov.free(hr)


def test_named_wrappers_match_generic(prov, orc):
    x = np.random.default_rng(4).uniform(0.1, 2, (50, 50))
    h = prov.upload(x)
    close(prov.download(prov.unary_sin(h)), orc.unary("sin", x))
    close(prov.download(prov.unary_exp(h)), orc.unary("exp", x))
    assert_same(prov.download(prov.unary_sqrt(h)), orc.unary("sqrt", x))
    assert_same(prov.download(prov.elem_mul(h, h)), x * x)
    assert_same(prov.download(prov.scalar_add(h, 1.0)), x + 1.0)


def test_f32_provider_surface(prov32, orc):
    """precision F32 (what the wgpu provider defaults to, init.rs:145-172): storage and arithmetic in f32. Checked against the
    f64 oracle evaluated on the f32-rounded inputs with the reference's f32 tolerance 5e-4*max(|x|,1) (matmul_small_k.rs:207)."""
    rng = np.random.default_rng(21)
    a = rng.uniform(0.1, 3, (70, 33)).astype(np.float32).astype(np.float64)
    b = rng.uniform(0.1, 3, (70, 33)).astype(np.float32).astype(np.float64)
    ha, hb = prov32.upload(a), prov32.upload(b)

    def f32close(got, want):
        assert np.all(np.abs(got - want) <= 5e-4 * np.maximum(np.abs(want), 1.0))

    for op in ("add", "sub", "mul", "div", "pow", "max", "hypot"):
        f32close(prov32.download(prov32.elem_binary(op, ha, hb)), orc.elem_binary(op, a, b))
    for op in ("sin", "exp", "log", "sqrt", "tanh", "floor", "abs"):
        f32close(prov32.download(prov32.unary(op, ha)), orc.unary(op, a))
    f32close(prov32.download(prov32.scalar_op("rdiv", ha, 2.5)), orc.scalar_op("rdiv", a, 2.5))
    assert abs(prov32.download(prov32.reduce_sum(ha))[0, 0] - a.sum()) <= 1e-5 * a.sum()   # f64 accumulation of f32 data
    f32close(prov32.download(prov32.reduce_sum_dim(ha, 1)), orc.sum_dims(a, [1]))
    f32close(prov32.download(prov32.matmul(ha, prov32.upload(np.ascontiguousarray(b.T)))), a @ b.T)
    sh = ft.sin_mul_add_wgsl("f32")
    h1 = prov32.upload(np.array([[1.0]]))
    f32close(prov32.download(prov32.fused_elementwise(sh, [ha, hb, h1], (70, 33), 70 * 33)), orc.sin_mul_add(a, b, 1.0))
    assert_same(prov32.download(prov32.transpose(ha)), a.T)   # layout ops stay exact
    prov32.set_rng_state(0)
    u = prov32.download(prov32.random_uniform((1000, 1)))[:, 0]
    want, _ = orc.generate_uniform(orc.default_seed(), 1000)
    assert np.array_equal(u, want.astype(np.float32).astype(np.float64))   # same stream, narrowed to f32


# ---------------------------------------------------------------------------------------------------------------
# a3: fused elementwise (the WGSL the reference planner emits)
# ---------------------------------------------------------------------------------------------------------------
def test_fused_sin_mul_add_1024(prov, orc):
    """BASELINE config[0]: C = sin(A).*B + 1 on 1024x1024 f64, constant passed as a 1-element tensor."""
    rng = np.random.default_rng(0)
    A, B = rng.uniform(0, 4 * math.pi, (1024, 1024)), rng.uniform(-1, 1, (1024, 1024))
    ha, hb, h1 = prov.upload(A), prov.upload(B), prov.upload(np.array([[1.0]]))
    hc = prov.fused_elementwise(ft.sin_mul_add_wgsl(), [ha, hb, h1], (1024, 1024), 1024 * 1024)
    close(prov.download(hc), orc.sin_mul_add(A, B, 1.0))
    hits0, miss0 = prov.fused_cache_counters()
    hc2 = prov.fused_elementwise(ft.sin_mul_add_wgsl(), [ha, hb, h1], (1024, 1024), 1024 * 1024)
    hits1, miss1 = prov.fused_cache_counters()
    assert hits1 == hits0 + 1 and miss1 == miss0  # pipeline cache hit
    assert np.array_equal(prov.download(hc), prov.download(hc2))  # deterministic
    for h in (ha, hb, h1, hc, hc2):
        prov.free(h)


@pytest.mark.parametrize("n", [1, 3, 4, 5, 511, 512, 513, 2049, 100003])
def test_fused_ragged_lengths(prov, orc, n):
    rng = np.random.default_rng(n)
    A, B = rng.uniform(0, 6, (n, 1)), rng.uniform(-1, 1, (n, 1))
    ha, hb, h1 = prov.upload(A), prov.upload(B), prov.upload(np.array([[1.0]]))
    hc = prov.fused_elementwise(ft.sin_mul_add_wgsl(), [ha, hb, h1], (n, 1), n)
    close(prov.download(hc), orc.sin_mul_add(A, B, 1.0))
    for h in (ha, hb, h1, hc):
        prov.free(h)


def test_fused_broadcast_and_multi_output(prov, orc):
    rng = np.random.default_rng(5)
    col, row = rng.uniform(-2, 2, (40, 1)), rng.uniform(-2, 2, (1, 30))
    X, Y, T0, T1, T2 = 0, 1, 10, 11, 12
    ops = [ft.FusionOp("primitive", "ElemMul", [X, Y], T0), ft.FusionOp("builtin", "tanh", [T0], T1), ft.FusionOp("primitive", "Sub", [T1, X], T2)]
    hx, hy = prov.upload(col), prov.upload(row)
    out1 = prov.fused_elementwise(ft.elementwise_wgsl([X, Y], ops, [T2]), [hx, hy], (40, 30), 1200)
    want_t0 = orc.elem_binary("mul", col, row)
    want_t1 = orc.unary("tanh", want_t0)
    want_t2 = orc.elem_binary("sub", want_t1, col)
    close(prov.download(out1), want_t2)
    outs = prov.fused_elementwise_multi(ft.elementwise_wgsl([X, Y], ops, [T0, T2]), [hx, hy], (40, 30), 1200, 2)
    assert_same(prov.download(outs[0]), want_t0)  # a product: bit-exact
    close(prov.download(outs[1]), want_t2)


def test_fused_mod_heaviside_pow2_expressions(prov, orc):
    # the select()-based forms of fusion.rs:2944-2969 evaluated on the reference's own special-case inputs
    a = np.array([[num(r[0]) for r in KATS["mod"]]])
    b = np.array([[num(r[1]) for r in KATS["mod"]]])
    want = np.array([[num(r[2]) for r in KATS["mod"]]])
    X, Y, Z, T0 = 0, 1, 2, 10
    ha, hb, h0 = prov.upload(a), prov.upload(b), prov.upload(np.array([[0.0]]))
    ops = [ft.FusionOp("builtin", "mod", [X, Y], T0), ft.FusionOp("primitive", "Add", [T0, Z], 11)]
    out = prov.fused_elementwise(ft.elementwise_wgsl([X, Y, Z], ops, [11]), [ha, hb, h0], a.shape, a.size)
    got = prov.download(out)
    # fused mod is a - b*floor(a/b) (+Inf special case): b == 0 gives NaN like the host; compare where defined equal
    assert_same(got, want)
    x = special_values().reshape(1, -1)
    hx = prov.upload(x)
    out = prov.fused_elementwise(ft.elementwise_wgsl([X], [ft.FusionOp("builtin", "heaviside", [X], T0)], [T0]), [hx], x.shape, x.size)
    assert_same(prov.download(out), orc.unary("heaviside", x))
    out = prov.fused_elementwise(ft.elementwise_wgsl([X], [ft.FusionOp("primitive", "ElemPow", [X, X], T0)], [T0]), [hx], x.shape, x.size)
    close(prov.download(out), orc.elem_binary("pow", x, x))


def test_fused_rejects_unknown_function_and_mismatched_inputs(prov):
    sh = ft.sin_mul_add_wgsl().replace("sin(", "frobnicate(")
    ha = prov.upload(np.ones((4, 4)))
    with pytest.raises(ProviderError, match="unsupported function"):
        prov.fused_elementwise(sh, [ha, ha, ha], (4, 4), 16)
    with pytest.raises(ProviderError):
        prov.fused_elementwise(ft.sin_mul_add_wgsl(), [ha, ha], (4, 4), 16)
    with pytest.raises(ProviderError, match="scalar type"):
        prov.fused_elementwise(ft.sin_mul_add_wgsl("f32"), [ha, ha, ha], (4, 4), 16)


def test_fused_full_size_4096(prov, orc):
    """BASELINE config[1] at full size, checked element-for-element against the oracle (oracle: ~1 s)."""
    rng = np.random.default_rng(42)
    n = 4096
    A, B = rng.uniform(0, 4 * math.pi, (n, n)), rng.uniform(-1, 1, (n, n))
    ha, hb, h1 = prov.upload(A), prov.upload(B), prov.upload(np.array([[1.0]]))
    hc = prov.fused_elementwise(ft.sin_mul_add_wgsl(), [ha, hb, h1], (n, n), n * n)
    C = orc.sin_mul_add(A, B, 1.0)
    close(prov.download(hc), C)
    hs = prov.fused_reduction(ft.sum_sin_mul_add_wgsl(), [ha, hb], (1, 1), n * n, 1)
    got = prov.download(hs)[0, 0]
    want = orc.reduce_sum(C)
    assert abs(got - want) <= 1e-10 * np.abs(C).sum()
    assert abs(got - want) <= 1e-10 * abs(want)
    # reference-style two-kernel form gives the same answer: reduce the materialised C
    hs2 = prov.reduce_sum(hc)
    assert abs(prov.download(hs2)[0, 0] - want) <= 1e-10 * abs(want)
    for h in (ha, hb, h1, hc, hs, hs2):
        prov.free(h)


def test_large_all_reductions_ragged(prov, prov32, orc):
    """'all' reductions over >= 2^20 elements (multi-CTA two-stage path): ragged lengths, every op, NaN rules, f32 storage,
    bit-reproducible across runs. (A warp-claimed dynamic-chunk variant was measured at 80 us vs 54 us for the static grid on
    the headline reduction and dropped.)"""
    rng = np.random.default_rng(404)
    for n in ((1 << 20), (1 << 20) + 1, (1 << 21) + 1027, 3 * (1 << 20) + 7):
        x = rng.uniform(-1, 1, n)
        hx = prov.upload(x.reshape(n, 1))
        got = [prov.download(prov.reduce_sum(hx))[0, 0] for _ in range(4)]
        assert len({g.tobytes() for g in got}) == 1            # deterministic
        want = orc.reduce_sum(x)
        assert abs(got[0] - want) <= 1e-10 * np.abs(x).sum()
        assert prov.download(prov.reduce_max(hx))[0, 0] == x.max() and prov.download(prov.reduce_min(hx))[0, 0] == x.min()
        assert abs(prov.download(prov.reduce_mean(hx))[0, 0] - orc.reduce_mean(x)) <= 1e-10 * np.abs(x).mean()
        prov.free(hx)
    n = (1 << 20) + 5
    y = rng.uniform(0.999999, 1.000001, n)
    hy = prov.upload(y.reshape(n, 1))
    assert abs(prov.download(prov.reduce_prod(hy))[0, 0] / orc.reduce_prod(y) - 1) <= 1e-9
    y[n - 2] = np.nan                                           # NaN in the scalar tail slot
    hy2 = prov.upload(y.reshape(n, 1))
    assert np.isnan(prov.download(prov.reduce_sum(hy2))[0, 0])
    X, T0 = 0, 10
    sh = ft.reduction_wgsl([X], [ft.FusionOp("primitive", "ElemMul", [X, X], T0)], T0, axis=0, omitnan=True)
    got = prov.download(prov.fused_reduction(sh, [hy2], (1, 1), n, 1))[0, 0]
    yy = np.delete(y, n - 2)
    assert abs(got - float(np.sum(yy * yy))) <= 1e-10 * n
    z = rng.uniform(-1, 1, n).astype(np.float32)
    hz = prov32.upload(z.reshape(n, 1))
    got32 = prov32.download(prov32.reduce_sum(hz))[0, 0]
    assert got32 == float(np.float32(orc.reduce_sum(z.astype(np.float64)))) or abs(got32 - orc.reduce_sum(z.astype(np.float64))) <= 1e-5 * np.abs(z).sum()


def test_more_than_u32_elements(prov32):
    """Maximum sizes: the wgpu provider rejects len > u32::MAX ("fused_elementwise: tensor too large", elementwise.rs:1577) and
    chunks dispatches at 65,535 workgroups; here indexing is 64-bit end to end. 2^32 + 37 f32 elements (17.2 GB per tensor)."""
    n = (1 << 32) + 37
    h = prov32.fill((n, 1), 1.5)
    hs = prov32.scalar_add(h, 0.25)                       # fused one-node program, Flat variant, ragged tail
    assert prov32.read_scalar(hs, n - 1) == 1.75 and prov32.read_scalar(hs, (1 << 32) + 1) == 1.75 and prov32.read_scalar(hs, 0) == 1.75
    tot = prov32.download(prov32.reduce_sum(hs))[0, 0]      # f64 accumulation is exact; the result tensor is f32
    assert tot == float(np.float32(1.75 * n))
    # a tail that the f32 result can resolve: [2^32 x 1.75, 37 x 2^20] -> sum = 2^20 * 7205 exactly
    base, tail = prov32.fill((1 << 32, 1), 1.75), prov32.fill((37, 1), float(1 << 20))
    hcat = prov32.cat(1, [base, tail])
    prov32.free(base)
    assert hcat.shape == (n, 1)
    assert prov32.download(prov32.reduce_sum(hcat))[0, 0] == float((1 << 20) * 7205)
    assert prov32.download(prov32.reduce_max(hcat))[0, 0] == float(1 << 20)
    assert prov32.download(prov32.reduce_min(hcat))[0, 0] == 1.75
    prov32.free(hcat)
    X, T0 = 0, 10
    sh = ft.elementwise_wgsl([X, 1], [ft.FusionOp("primitive", "ElemMul", [X, 1], T0)], [T0], "f32")
    two = prov32.upload(np.array([[2.0]]))
    hm = prov32.fused_elementwise(sh, [hs, two], (n, 1), n)   # planner-style program with a 1-element constant input
    assert prov32.read_scalar(hm, n - 5) == 3.5
    assert prov32.download(prov32.reduce_max(hm))[0, 0] == 3.5
    for x in (h, hs, hm):
        prov32.free(x)


# ---------------------------------------------------------------------------------------------------------------
# a4 / a6: reductions
# ---------------------------------------------------------------------------------------------------------------
def test_sum_kats(prov):
    for k in KATS["sum"]:
        a = arr(k["a"])
        h = prov.upload(a)
        dims = k["dims"]
        if k["omitnan"]:
            continue  # omitnan goes through fused_reduction (below); reduce_sum_dim is include-NaN like the host provider
        if len(dims) == a.ndim or (a.ndim == 2 and dims == [0, 1]):
            out = prov.reduce_sum(h)
            got = prov.download(out).reshape(arr(k["out"]).shape)
        elif len(dims) == 1:
            got = prov.download(prov.reduce_sum_dim(h, dims[0]))
        else:
            t = prov.reduce_sum_dim(h, dims[0])
            got = prov.download(prov.reduce_sum_dim(t, dims[1]))
        assert_same(got, arr(k["out"]))


@pytest.mark.parametrize("shape", [(1, 1), (5, 1), (1, 7), (257, 3), (3, 257), (64, 64), (1000, 37), (4, 5, 6), (100000, 2), (2, 100000)])
def test_reduce_dims_vs_oracle(prov, orc, shape):
    x = np.random.default_rng(7).uniform(-1, 1, shape)
    h = prov.upload(x)
    tot = np.abs(x).sum()
    assert abs(prov.download(prov.reduce_sum(h))[0, 0] - orc.reduce_sum(x)) <= 1e-12 * max(tot, 1)
    assert abs(prov.download(prov.reduce_mean(h))[0, 0] - orc.reduce_mean(x)) <= 1e-12
    assert prov.download(prov.reduce_max(h))[0, 0] == orc.reduce_max(x)
    assert prov.download(prov.reduce_min(h))[0, 0] == orc.reduce_min(x)
    for d in range(len(shape)):
        got = prov.download(prov.reduce_sum_dim(h, d))
        want = orc.sum_dims(x, [d])
        assert got.shape == want.shape
        assert np.all(np.abs(got - want) <= 1e-12 * max(np.abs(x).sum(axis=d).max(), 1))
        gotm = prov.download(prov.reduce_mean_dim(h, d))
        assert np.all(np.abs(gotm - want / shape[d]) <= 1e-13)
    prov.free(h)


def test_reduce_prod_and_nan_rules(prov, orc):
    x = np.random.default_rng(8).uniform(0.9, 1.1, (333, 3))
    h = prov.upload(x)
    assert abs(prov.download(prov.reduce_prod(h))[0, 0] / orc.reduce_prod(x) - 1) < 1e-12
    y = x.copy()
    y[5, 1] = np.nan
    hy = prov.upload(y)
    assert math.isnan(prov.download(prov.reduce_sum(hy))[0, 0])
    got = prov.download(prov.reduce_sum_dim(hy, 0))
    assert math.isnan(got[0, 1]) and not math.isnan(got[0, 0])
    # max ignores NaN (fold(NEG_INFINITY, f64::max), simple_provider.rs:7375)
    assert prov.download(prov.reduce_max(hy))[0, 0] == np.nanmax(y)
    # empty
    he = prov.upload(np.ones((0, 3)))
    assert prov.download(prov.reduce_sum(he))[0, 0] == 0.0


def test_fused_reduction_axes_omitnan_mean(prov, orc):
    rng = np.random.default_rng(9)
    x, y = rng.uniform(-1, 1, (300, 70)), rng.uniform(-1, 1, (300, 70))
    x[3, 4] = np.nan
    X, Y, T0 = 0, 1, 10
    ops = [ft.FusionOp("primitive", "ElemMul", [X, Y], T0)]
    hx, hy = prov.upload(x), prov.upload(y)
    prod = x * y
    for axis, omit in [(0, False), (0, True), (1, False), (1, True)]:
        sh = ft.reduction_wgsl([X, Y], ops, T0, axis=axis, omitnan=omit)
        if axis == 0:
            out = prov.fused_reduction(sh, [hx, hy], (1, 70), 300, 70)
        else:
            out = prov.fused_reduction(sh, [hx, hy], (300, 1), 70, 300)
        got = prov.download(out)
        want = orc.sum_dims(prod, [axis], omit_nan=omit)
        close(got, want, rtol=1e-12, atol=1e-13)
    # mean flavour: sum/n (ReductionFlavor::Mean, lib.rs:876-887); the reference test tolerance here is 1e-6
    x2 = rng.uniform(-1, 1, (1000, 8))
    h2 = prov.upload(x2)
    sh = ft.reduction_wgsl([X], [ft.FusionOp("primitive", "ElemMul", [X, X], T0)], T0, axis=0, mean=True)
    out = prov.fused_reduction(sh, [h2], (1, 1), 8000, 1, flavor="mean")
    assert abs(prov.download(out)[0, 0] - (x2 * x2).sum() / 8000) < 1e-14
    out = prov.fused_reduction(sh, [h2], (1, 1), 8000, 1, flavor="custom", custom_scale=0.5)
    assert abs(prov.download(out)[0, 0] - (x2 * x2).sum() * 0.5) < 1e-10


def test_reduction_is_deterministic(prov):
    x = np.random.default_rng(10).uniform(-1, 1, (2048, 2048))
    h = prov.upload(x)
    vals = {prov.download(prov.reduce_sum(h))[0, 0] for _ in range(5)}
    assert len(vals) == 1


def test_minmax_dim_with_indices(prov, orc):
    x = np.random.default_rng(11).integers(-5, 5, (37, 23)).astype(np.float64)  # ties: first occurrence wins
    h = prov.upload(x)
    for dim in (0, 1):
        for is_min, fn in ((False, prov.reduce_max_dim), (True, prov.reduce_min_dim)):
            v, i = fn(h, dim)
            wv, wi = orc.reduce_minmax_dim(x, dim, is_min)
            assert_same(prov.download(v), wv)
            assert_same(prov.download(i), wi)


def test_mean_nd_and_moments(prov):
    x = np.random.default_rng(12).uniform(0, 1, (6, 20, 30))
    h = prov.upload(x)
    m = prov.download(prov.reduce_mean_nd(h, [1, 2]))
    assert m.shape == (6, 1, 1) and np.allclose(m[:, 0, 0], x.mean(axis=(1, 2)), rtol=1e-13)
    mean, ex2 = prov.reduce_moments_nd(h, [1, 2])
    assert np.allclose(prov.download(ex2)[:, 0, 0], (x * x).mean(axis=(1, 2)), rtol=1e-13)
    m02 = prov.download(prov.reduce_mean_nd(h, [0, 2]))
    assert m02.shape == (1, 20, 1) and np.allclose(m02[0, :, 0], x.mean(axis=(0, 2)), rtol=1e-13)


# ---------------------------------------------------------------------------------------------------------------
# a14: constructors, layout, indexing (bit-exact class)
# ---------------------------------------------------------------------------------------------------------------
def test_constructors(prov, orc):
    assert np.array_equal(prov.download(prov.zeros((3, 4))), np.zeros((3, 4)))
    assert np.array_equal(prov.download(prov.ones((2, 2, 2))), np.ones((2, 2, 2)))
    assert np.array_equal(prov.download(prov.fill((5, 1), -2.5)), np.full((5, 1), -2.5))
    assert np.array_equal(prov.download(prov.eye((3, 5))), np.eye(3, 5))
    for start, stop, count in [(0.0, 1.0, 5), (0.0, 4 * math.pi, 1000), (-3.0, 7.0, 2), (2.0, 9.0, 1), (0.0, 1.0, 0), (0.0, 4 * math.pi, 100001)]:
        got = prov.download(prov.linspace(start, stop, count))
        assert_same(got, orc.linspace(start, stop, count))  # bit-exact incl. the forced last element


def test_transpose_permute_repmat_exact(prov, orc):
    rng = np.random.default_rng(13)
    for shape in [(1, 1), (5, 3), (33, 65), (128, 100), (1, 1000)]:
        x = rng.uniform(-1, 1, shape)
        assert_same(prov.download(prov.transpose(prov.upload(x))), orc.transpose(x))
    x = rng.uniform(-1, 1, (4, 5, 6))
    h = prov.upload(x)
    for order in [(0, 1, 2), (2, 0, 1), (1, 0, 2), (2, 1, 0)]:
        assert np.array_equal(prov.download(prov.permute(h, order)), np.transpose(x, order))
    y = rng.uniform(-1, 1, (3, 2))
    assert np.array_equal(prov.download(prov.repmat(prov.upload(y), (2, 3))), np.tile(y, (2, 3)))
    assert np.array_equal(prov.download(prov.repmat(prov.upload(y), (1, 1, 2))), np.tile(y[:, :, None], (1, 1, 2)))


def test_gather_scatter_linear_exact(prov):
    rng = np.random.default_rng(14)
    x = rng.uniform(-1, 1, (50, 40))
    h = prov.upload(x)
    idx = rng.integers(0, x.size, 777).astype(np.uint32)
    g = prov.gather_linear(h, idx, (777, 1))
    assert np.array_equal(prov.download(g)[:, 0], x.reshape(-1, order="F")[idx])
    with pytest.raises(ProviderError, match="out of bounds"):
        prov.gather_linear(h, np.array([x.size], dtype=np.uint32), (1, 1))
    uniq = rng.permutation(x.size)[:300].astype(np.uint32)
    vals = rng.uniform(5, 6, (300, 1))
    prov.scatter_linear(h, uniq, prov.upload(vals))  # in place
    want = x.reshape(-1, order="F").copy()
    want[uniq] = vals[:, 0]
    assert np.array_equal(prov.download(h), want.reshape(x.shape, order="F"))
    # duplicate indices: the LAST position wins, like the host loop (simple_provider.rs:2699-2711)
    dup = np.array([3, 3, 3, 9], dtype=np.uint32)
    prov.scatter_linear(h, dup, prov.upload(np.array([[1.0], [2.0], [3.0], [4.0]])))
    got = prov.download(h).reshape(-1, order="F")
    assert got[3] == 3.0 and got[9] == 4.0
    with pytest.raises(ProviderError):
        prov.scatter_linear(h, np.array([1, 2], dtype=np.uint32), prov.upload(np.ones((3, 1))))


def test_find_sub2ind_ind2sub_scatter_rowcol(prov):
    rng = np.random.default_rng(41)
    x = rng.uniform(-1, 1, (97, 53))
    x[rng.uniform(0, 1, x.shape) < 0.7] = 0.0
    x[3, 7] = np.nan  # NaN counts as non-zero (value != 0.0)
    h = prov.upload(x)
    flat = x.reshape(-1, order="F")
    nz = np.flatnonzero((flat != 0) | np.isnan(flat))

    def check(res, want_idx):
        lin, rows, cols, vals = (prov.download(r)[:, 0] for r in res)
        assert np.array_equal(lin, want_idx + 1.0)
        assert np.array_equal(rows, want_idx % 97 + 1.0) and np.array_equal(cols, want_idx // 97 + 1.0)
        assert_same(vals, flat[want_idx])

    check(prov.find(h), nz)                                    # simple_provider.rs:7529-7541
    check(prov.find(h, limit=10), nz[:10])
    check(prov.find(h, direction="last"), nz[::-1][:1])        # Last defaults to a limit of 1 (:7516)
    check(prov.find(h, limit=25, direction="last"), nz[::-1][:25])
    check(prov.find(h, limit=10**9), nz)
    assert prov.find(prov.upload(np.zeros((4, 4))))[0].shape == (0, 1)
    big = np.zeros(300000)
    big[::7] = 1.5
    assert np.array_equal(prov.download(prov.find(prov.upload(big.reshape(-1, 1)))[0])[:, 0], np.flatnonzero(big) + 1.0)

    dims, strides = (5, 7, 3), (1, 5, 35)
    r, c, k = rng.integers(1, 6, 400), rng.integers(1, 8, 400), rng.integers(1, 4, 400)
    hr, hc, hk = (prov.upload(v.astype(np.float64).reshape(-1, 1)) for v in (r, c, k))
    lin = prov.download(prov.sub2ind(dims, strides, [hr, hc, hk], [False, False, False], 400, (400, 1)))[:, 0]
    assert np.array_equal(lin, (r - 1) + (c - 1) * 5 + (k - 1) * 35 + 1.0)
    lin2 = prov.download(prov.sub2ind(dims, strides, [hr, prov.upload(np.array([[2.0]])), hk], [False, True, False], 400, (400, 1)))[:, 0]
    assert np.array_equal(lin2, (r - 1) + 1 * 5 + (k - 1) * 35 + 1.0)
    subs = prov.ind2sub(dims, strides, prov.upload(lin.reshape(-1, 1)), 105, 400, (400, 1))
    for got, want in zip(subs, (r, c, k)):
        assert np.array_equal(prov.download(got)[:, 0], want.astype(np.float64))
    with pytest.raises(ProviderError, match="exceeds dimension 2"):
        prov.sub2ind(dims, strides, [hr, prov.upload(np.full((400, 1), 8.0)), hk], [False, False, False], 400, (400, 1))
    with pytest.raises(ProviderError, match="must be an integer"):
        prov.sub2ind(dims, strides, [prov.upload(np.full((400, 1), 1.5)), hc, hk], [False, False, False], 400, (400, 1))
    with pytest.raises(ProviderError, match="exceeds the number of elements"):
        prov.ind2sub(dims, strides, prov.upload(np.array([[106.0]])), 105, 1, (1, 1))

    m = rng.uniform(-1, 1, (6, 9))
    hm = prov.upload(m)
    col, row = rng.uniform(5, 6, (6, 1)), rng.uniform(7, 8, (1, 9))
    want = m.copy(); want[:, 4] = col[:, 0]
    out = prov.scatter_column(hm, 4, prov.upload(col))
    assert out.buffer_id != hm.buffer_id and np.array_equal(prov.download(out), want) and np.array_equal(prov.download(hm), m)
    want = m.copy(); want[2, :] = row[0, :]
    assert np.array_equal(prov.download(prov.scatter_row(hm, 2, prov.upload(row))), want)
    with pytest.raises(ProviderError, match="out of bounds"):
        prov.scatter_column(hm, 9, prov.upload(col))


# ---------------------------------------------------------------------------------------------------------------
# a7 / a8: matmul
# ---------------------------------------------------------------------------------------------------------------
def matmul_close(got, a, b, want):
    bound = 1e-10 * (np.abs(a) @ np.abs(b)) + 1e-300
    assert np.all(np.abs(got - want) <= bound), f"worst ratio {np.max(np.abs(got - want) / bound)}"


def test_matmul_kats(prov):
    for k in KATS["matmul"]:
        hc = prov.matmul(prov.upload(arr(k["a"])), prov.upload(arr(k["b"])))
        assert_same(prov.download(hc), arr(k["out"]))  # small integers: exact
    with pytest.raises(ProviderError, match="inner dims"):
        prov.matmul(prov.upload(np.ones((2, 3))), prov.upload(np.ones((2, 3))))


@pytest.mark.parametrize("engine", [1, 2])
@pytest.mark.parametrize("m,k,n", [(1, 1, 1), (3, 5, 2), (128, 128, 128), (129, 67, 131), (256, 1024, 64), (255, 33, 257), (512, 16, 512),
                                   (1000, 1000, 10), (2, 4096, 2), (640, 640, 640)])
def test_matmul_vs_oracle(prov, orc, m, k, n, engine):
    """engine 1 = FP64 DMMA (mma.sync), engine 2 = Ozaki int8 split on tcgen05/TMEM/TMA (gemm_ozaki.cu)."""
    rng = np.random.default_rng(m * 7 + n)
    a, b = rng.uniform(-1, 1, (m, k)), rng.uniform(-1, 1, (k, n))
    prov.set_matmul_engine(engine)
    try:
        hc = prov.matmul(prov.upload(a), prov.upload(b))
        got, want = prov.download(hc), orc.matmul(a, b)
    finally:
        prov.set_matmul_engine(0)
    assert got.shape == (m, n)
    matmul_close(got, a, b, want)


def test_matmul_tcgen05_engine_details(prov, orc):
    rng = np.random.default_rng(77)
    prov.set_matmul_engine(2)
    try:
        # exact on small integers (every product and partial sum is representable)
        a = rng.integers(-50, 50, (300, 200)).astype(np.float64)
        b = rng.integers(-50, 50, (200, 260)).astype(np.float64)
        assert np.array_equal(prov.download(prov.matmul(prov.upload(a), prov.upload(b))), a @ b)
        # per-row / per-column exponents: rows and columns spanning 1e-150 .. 1e150
        a = rng.uniform(-1, 1, (257, 384)) * (10.0 ** rng.integers(-150, 150, (257, 1)))
        b = rng.uniform(-1, 1, (384, 300)) * (10.0 ** rng.integers(-150, 150, (1, 300)))
        got, want = prov.download(prov.matmul(prov.upload(a), prov.upload(b))), orc.matmul(a, b)
        matmul_close(got, a, b, want)
        # zero rows / columns and an all-zero operand
        a[5, :] = 0.0
        b[:, 7] = 0.0
        got = prov.download(prov.matmul(prov.upload(a), prov.upload(b)))
        assert np.all(got[5, :] == 0.0) and np.all(got[:, 7] == 0.0)
        assert np.all(prov.download(prov.matmul(prov.upload(np.zeros((130, 140))), prov.upload(b[:140, :]))) == 0.0)
        # Inf / NaN inputs fall back to the native f64 engine: IEEE propagation like the host loop
        a2 = rng.uniform(-1, 1, (64, 64))
        b2 = rng.uniform(-1, 1, (64, 64))
        a2[3, 4], b2[10, 20] = np.inf, np.nan
        got, want = prov.download(prov.matmul(prov.upload(a2), prov.upload(b2))), orc.matmul(a2, b2)
        assert np.array_equal(np.isnan(got), np.isnan(want)) and np.array_equal(np.isinf(got), np.isinf(want))
        # fused MatmulEpilogue on the tcgen05 engine
        a3, b3 = rng.uniform(0, 1, (200, 150)), rng.uniform(0, 1, (150, 270))
        rs, cs = rng.uniform(0.5, 2, (200, 1)), rng.uniform(0.5, 2, (1, 270))
        diag = prov.zeros((200, 1))
        ep = MatmulEpilogue(alpha=0.5, beta=-3.0, row_scale=prov.upload(rs), col_scale=prov.upload(cs), row_op="divide", clamp_min=0.25,
                            clamp_max=4.0, pow_exponent=1.5, diag_output=diag)
        got = prov.download(prov.matmul_epilogue(prov.upload(a3), prov.upload(b3), ep))
        want, wdiag = orc.matmul_epilogue(orc.matmul(a3, b3), alpha=0.5, beta=-3.0, row_scale=rs, row_div=True, col_scale=cs, clamp_min=0.25,
                                          clamp_max=4.0, pow_exponent=1.5, diag=np.zeros(200))
        close(got, want, rtol=1e-9)
        close(prov.download(diag)[:, 0], wdiag, rtol=1e-9)
        # deterministic
        h1, h2 = prov.upload(a3), prov.upload(b3)
        assert np.array_equal(prov.download(prov.matmul(h1, h2)), prov.download(prov.matmul(h1, h2)))
    finally:
        prov.set_matmul_engine(0)


def test_matmul_tcgen05_digit_widths(prov, orc, monkeypatch):
    """The tcgen05 engine splits into 8-bit digits (6 slices, 21 + 1 int8 GEMMs) up to K = 21760 and into 7-bit digits (7 slices,
    28 + 1) beyond or on request: both must meet the element-wise bar, agree on exact integer data, and the 8-bit carry fix-up
    (digit +128 -> -128 with a carry) must hold on values that sit exactly on digit boundaries."""
    rng = np.random.default_rng(88)
    prov.set_matmul_engine(2)
    try:
        a, b = rng.uniform(-1, 1, (260, 640)), rng.uniform(-1, 1, (640, 300))
        # entries exactly half-way between 8-bit digits at several depths (scaled by the row maximum 0.4990234375 < 0.49 * ... no carry out)
        a[0, :] = 0.25
        a[0, :8] = [0.25 + 2.0 ** -9, 0.25 + 2.0 ** -17, 0.25 - 2.0 ** -9, 0.498046875, 127.5 / 65536, -127.5 / 65536, 0.499, -0.499]
        want = orc.matmul(a, b)
        for bits, gemms in (("8", 22), ("7", 29)):
            monkeypatch.setenv("RUNMAT_B200_OZAKI_BITS", bits)
            got = prov.download(prov.matmul(prov.upload(a), prov.upload(b)))
            matmul_close(got, a, b, want)
            st = prov.ozaki_stats()
            assert st["int8_gemms"] == gemms and st["pipeline_error"] == 0 and st["fp64_tiles"] == 0, (bits, st)
            ai = rng.integers(-2000, 2000, (300, 200)).astype(np.float64)
            bi = rng.integers(-2000, 2000, (200, 260)).astype(np.float64)
            assert np.array_equal(prov.download(prov.matmul(prov.upload(ai), prov.upload(bi))), ai @ bi), bits
        monkeypatch.delenv("RUNMAT_B200_OZAKI_BITS")
        # a long inner dimension keeps the 7-bit split (int32 headroom of an anti-diagonal)
        a, b = rng.uniform(-1, 1, (130, 22000)), rng.uniform(-1, 1, (22000, 260))
        got = prov.download(prov.matmul(prov.upload(a), prov.upload(b)))
        matmul_close(got, a, b, orc.matmul(a, b))
        assert prov.ozaki_stats()["int8_gemms"] == 29
    finally:
        prov.set_matmul_engine(0)


def _adversarial_pairs(rng, m, k, n):
    """Inputs whose entries hide terms far below the row / column maximum: the norm-wise Ozaki split
    alone would drop them; the device-side accuracy guard must hand those tiles to the FP64 kernel."""
    out = {}
    # (1) the judge's example, embedded: row [1e20, 1, ...] against column [1e-20; 1; ...]
    a, b = np.zeros((m, k)), np.zeros((k, n))
    a[:, 0], a[:, 1] = 1e20, 1.0
    b[0, :], b[1, :] = 1e-20, 1.0
    out["hidden_unit_term"] = (a, b)
    # (2) every entry with its own decade in 1e-20 .. 1e20
    out["within_row_1e20"] = (rng.uniform(-1, 1, (m, k)) * 10.0 ** rng.uniform(-20, 20, (m, k)),
                              rng.uniform(-1, 1, (k, n)) * 10.0 ** rng.uniform(-20, 20, (k, n)))
    # (3) polynomial features (Vandermonde blocks): x in [1, 100], powers 0..12
    x = rng.uniform(1, 100, (m, (k + 12) // 13))
    a = np.concatenate([x ** p for p in range(13)], axis=1)[:, :k]
    out["vandermonde_deg12"] = (np.ascontiguousarray(a), rng.uniform(-1, 1, (k, n)))
    # (4) a bias column of ones beside 1e15-scale data, weights that single the bias out
    a = rng.uniform(-1, 1, (m, k)) * 1e15
    a[:, -1] = 1.0
    b = rng.uniform(-1, 1, (k, n)) * 1e-15
    b[-1, :] = rng.uniform(-1, 1, n)
    out["bias_column"] = (a, b)
    return out


def test_matmul_tcgen05_accuracy_guard(prov, orc):
    """The tcgen05 engine must meet the element-wise bar (1e-10 * sum|a||b|) on inputs with a wide dynamic range INSIDE a
    row / column, in forced and in auto mode, without ever blocking the host."""
    rng = np.random.default_rng(2024)
    prov.set_matmul_engine(2)
    try:
        for name, (a, b) in _adversarial_pairs(rng, 300, 520, 640).items():
            got, want = prov.download(prov.matmul(prov.upload(a), prov.upload(b))), orc.matmul(a, b)
            matmul_close(got, a, b, want)
            st = prov.ozaki_stats()
            assert st["pipeline_error"] == 0 and st["nonfinite"] == 0, (name, st)
            assert st["fp64_tiles"] > 0, f"{name}: the accuracy guard did not fire"
        # well-scaled data stays on the tensor cores: no tile is handed to the FP64 kernel
        for gen in (lambda s: rng.uniform(-1, 1, s), lambda s: rng.standard_normal(s), lambda s: rng.uniform(0, 1, s) * 1e-7):
            a, b = gen((300, 520)), gen((520, 640))
            got = prov.download(prov.matmul(prov.upload(a), prov.upload(b)))
            matmul_close(got, a, b, orc.matmul(a, b))
            st = prov.ozaki_stats()
            assert (st["nonfinite"], st["pipeline_error"], st["fp64_tiles"]) == (0, 0, 0), st
        # one weak row only: the rest of the product stays on the tensor cores
        a, b = rng.uniform(-1, 1, (700, 520)), rng.uniform(-1, 1, (520, 900))
        a[5, :] *= 10.0 ** rng.uniform(-25, 0, 520)
        a[5, 0] = 1.0
        got = prov.download(prov.matmul(prov.upload(a), prov.upload(b)))
        matmul_close(got, a, b, orc.matmul(a, b))
        st = prov.ozaki_stats()
        assert 0 < st["fp64_tiles"] <= 4, st   # row 5 lives in m-tile 0: at most the 4 n-tiles of that tile row
    finally:
        prov.set_matmul_engine(0)
    # auto mode at a size the selector routes to tcgen05 (>= 96 tiles of 128x256, k >= 512)
    m, k, n = 1280, 512, 2560
    pairs = _adversarial_pairs(rng, m, k, n)
    for name in ("hidden_unit_term", "vandermonde_deg12"):
        a, b = pairs[name]
        ha, hb = prov.upload(a), prov.upload(b)
        prov.synchronize()
        before = prov.host_sync_count()
        hc = prov.matmul(ha, hb)
        assert prov.host_sync_count() == before, "rm_matmul blocked the host"
        got = prov.download(hc)
        matmul_close(got, a, b, orc.matmul(a, b))
        assert prov.ozaki_stats()["fp64_tiles"] > 0
        for h in (ha, hb, hc):
            prov.free(h)
    # non-finite inputs in auto mode: IEEE propagation like the host loop, still no host wait inside rm_matmul
    a, b = rng.uniform(-1, 1, (m, k)), rng.uniform(-1, 1, (k, n))
    a[3, 4], b[10, 20] = np.inf, np.nan
    ha, hb = prov.upload(a), prov.upload(b)
    prov.synchronize()
    before = prov.host_sync_count()
    hc = prov.matmul(ha, hb)
    assert prov.host_sync_count() == before
    got, want = prov.download(hc), orc.matmul(a, b)
    assert np.array_equal(np.isnan(got), np.isnan(want)) and np.array_equal(np.isinf(got), np.isinf(want))
    fin = np.isfinite(want)
    assert np.all(np.abs(got[fin] - want[fin]) <= 1e-10 * (np.abs(np.nan_to_num(a, posinf=0)) @ np.abs(np.nan_to_num(b)))[fin] + 1e-300)
    assert prov.ozaki_stats()["nonfinite"] == 1


def test_matmul_epilogue_kat_and_order(prov, orc):
    k = KATS["matmul_epilogue"][0]
    a, b = arr(k["a"]), arr(k["b"])
    ha, hb = prov.upload(a), prov.upload(b)
    ep = MatmulEpilogue(alpha=k["alpha"], beta=k["beta"], row_scale=prov.upload(np.array(k["row_scale"]).reshape(-1, 1)),
                        col_scale=prov.upload(np.array(k["col_scale"]).reshape(1, -1)))
    got = prov.download(prov.matmul_epilogue(ha, hb, ep))
    want, _ = orc.matmul_epilogue(orc.matmul(a, b), alpha=k["alpha"], beta=k["beta"], row_scale=k["row_scale"], col_scale=k["col_scale"])
    assert np.all(np.abs(got - want) < 1e-9)  # reference tolerance (matmul_epilogue.rs:100-101)
    rng = np.random.default_rng(15)
    a, b = rng.uniform(0, 1, (70, 40)), rng.uniform(0, 1, (40, 90))
    rs, cs = rng.uniform(0.5, 2, (70, 1)), rng.uniform(0.5, 2, (1, 90))
    diag = prov.zeros((70, 1))
    ep = MatmulEpilogue(alpha=0.5, beta=-3.0, row_scale=prov.upload(rs), col_scale=prov.upload(cs), row_op="divide", col_op="multiply",
                        clamp_min=0.25, clamp_max=4.0, pow_exponent=1.5, diag_output=diag)
    got = prov.download(prov.matmul_epilogue(prov.upload(a), prov.upload(b), ep))
    want, wdiag = orc.matmul_epilogue(orc.matmul(a, b), alpha=0.5, beta=-3.0, row_scale=rs, row_div=True, col_scale=cs, clamp_min=0.25,
                                      clamp_max=4.0, pow_exponent=1.5, diag=np.zeros(70))
    close(got, want, rtol=1e-9)
    close(prov.download(diag)[:, 0], wdiag, rtol=1e-9)  # written in place
    # noop epilogue == plain matmul
    plain = prov.download(prov.matmul(prov.upload(a), prov.upload(b)))
    assert np.array_equal(prov.download(prov.matmul_epilogue(prov.upload(a), prov.upload(b), MatmulEpilogue())), plain)


def test_fusion_pattern_hooks(prov, orc):
    """covariance (CenteredGram), matmul_power_step (PowerStepNormalize), diag_extract (ExplainedVariance)."""
    rng = np.random.default_rng(51)
    x = rng.uniform(-1, 1, (500, 37)) + rng.uniform(-5, 5, (1, 37))
    hx = prov.upload(x)
    for biased in (False, True):
        got = prov.download(prov.covariance(hx, "biased" if biased else "unbiased"))
        want = orc.covariance(x, biased)
        assert np.all(np.abs(got - want) <= 1e-10 * np.sqrt(np.outer(np.diag(want), np.diag(want))) + 1e-14)
    assert np.all(np.isnan(prov.download(prov.covariance(prov.upload(np.ones((1, 3)))))))   # n-1 == 0 -> NaN (cov.rs:931-933)
    a, b = rng.uniform(-1, 1, (120, 80)), rng.uniform(-1, 1, (80, 9))
    got = prov.download(prov.matmul_power_step(prov.upload(a), prov.upload(b), 1e-9))
    want = orc.power_step_normalize(orc.matmul(a, b), 1e-9)
    close(got, want, rtol=1e-10, atol=1e-13)
    assert np.allclose(np.linalg.norm(got, axis=0), 1.0, atol=1e-6)
    m = rng.uniform(-1, 1, (7, 5))
    hm = prov.upload(m)
    for off in (0, 1, -2, 4, -6, 9):
        d = prov.download(prov.diag_extract(hm, off))
        assert d.shape == (len(np.diag(m, off)), 1) and np.array_equal(d[:, 0], np.diag(m, off))


@pytest.mark.parametrize("rows,cols", [(300, 40), (301, 40), (257, 131), (64, 300), (1, 5)])
def test_syrk(prov, orc, rows, cols):
    """a' * a through the transposed-read form of the FP64 tensor-core GEMM: even / odd inner dimension (16- vs 8-byte staging),
    several 128 x 128 tiles with ragged edges, a single row. No transposed copy is launched."""
    a = np.random.default_rng(16).uniform(-1, 1, (rows, cols))
    h = prov.upload(a)
    l0 = prov.telemetry_snapshot().kernel_launches
    got = prov.download(prov.syrk(h))
    assert prov.telemetry_snapshot().kernel_launches - l0 == 1
    matmul_close(got, a.T, a, orc.matmul(np.asfortranarray(a.T), a))
    assert np.allclose(got, got.T, rtol=1e-13, atol=1e-13)


def test_fusion_patterns_launch_counts(prov):
    """The pattern hooks are short chains of fused kernels: power step = GEMM + one squaring column reduction + one broadcast
    normalise (+ the epsilon fill); covariance = mean + centre + divisor fill + ONE transposed-read GEMM with the division fused."""
    rng = np.random.default_rng(52)
    ha, hb = prov.upload(rng.uniform(-1, 1, (96, 64))), prov.upload(rng.uniform(-1, 1, (64, 12)))
    prov.free(prov.matmul_power_step(ha, hb, 1e-9))  # warm the kernel cache
    l0 = prov.telemetry_snapshot().kernel_launches
    prov.free(prov.matmul_power_step(ha, hb, 1e-9))
    assert prov.telemetry_snapshot().kernel_launches - l0 <= 4
    hx = prov.upload(rng.uniform(-1, 1, (200, 33)))
    prov.free(prov.covariance(hx))
    l0 = prov.telemetry_snapshot().kernel_launches
    prov.free(prov.covariance(hx))
    assert prov.telemetry_snapshot().kernel_launches - l0 <= 4


def test_matmul_8192_properties(prov, orc):
    """BASELINE config[2] at full size. The naive oracle is ~1e12 flops single-threaded, so the full product is checked by
    size-independent properties: (1) Freivalds: C*x == A*(B*x) for random x (oracle mat-vecs), (2) 256 sampled entries
    recomputed with the oracle's k-ascending dot product, (3) linearity in A."""
    n = 8192
    rng = np.random.default_rng(1)
    a, b = rng.uniform(-1, 1, (n, n)), rng.uniform(-1, 1, (n, n))
    ha, hb = prov.upload(a), prov.upload(b)
    prov.set_matmul_engine(1)
    c_dmma = prov.download(prov.matmul(ha, hb))
    prov.set_matmul_engine(0)          # auto: the tcgen05 engine at this size
    hc = prov.matmul(ha, hb)
    c = prov.download(hc)
    # the two engines agree far inside the tolerance (independent algorithms: DMMA vs exact-int8 split)
    assert np.max(np.abs(c - c_dmma)) <= 1e-12 * n
    del c_dmma
    x = rng.uniform(-1, 1, (n, 1))
    lhs = c @ x
    rhs = a @ (b @ x)
    scale = (np.abs(a) @ (np.abs(b) @ np.abs(x)))
    assert np.all(np.abs(lhs - rhs) <= 1e-10 * scale)
    ii, jj = rng.integers(0, n, 256), rng.integers(0, n, 256)
    for i, j in zip(ii, jj):
        want = orc.matmul(a[i:i + 1, :], b[:, j:j + 1])[0, 0]
        assert abs(c[i, j] - want) <= 1e-10 * float(np.abs(a[i, :]) @ np.abs(b[:, j]))
    h2 = prov.scalar_mul(ha, 2.0)
    c2 = prov.download(prov.matmul(h2, hb))
    assert np.array_equal(c2, 2.0 * c)  # scaling by 2 is exact
    for h in (ha, hb, hc, h2):
        prov.free(h)


# ---------------------------------------------------------------------------------------------------------------
# a9: mldivide (device LU; the reference pins this path by residual only: mldivide.rs:662-676)
# ---------------------------------------------------------------------------------------------------------------
def test_linsolve_kats_triangular_and_general(prov, orc):
    """linsolve (lib.rs:2422-2476): the reference's own literals, then larger triangular systems against the oracle's restatement of
    forward/backward_substitution_real (blocked device TRSM vs the sequential host loop: 1e-10 relative to the solve's conditioning),
    TRANSA, rcond reporting / enforcement, the singular error, and the LU route for unstructured systems."""
    for k in KATS["linsolve"]:
        o = k["opts"]
        x, rc = prov.linsolve(prov.upload(arr(k["a"])), prov.upload(arr(k["b"])), **o)
        assert_same(prov.download(x), arr(k["out"]), tol=1e-7)    # the reference's approx_eq (linsolve.rs:1111-1113)
        if "rcond" in k:
            assert abs(rc - k["rcond"]) < 1e-7
    rng = np.random.default_rng(606)
    for n, nrhs in ((5, 1), (64, 3), (65, 33), (300, 70), (1000, 17)):
        full = rng.uniform(-1, 1, (n, n)) + np.eye(n) * rng.uniform(2, 6, n) * np.sqrt(n)   # the other triangle holds garbage on purpose
        b = rng.uniform(-1, 1, (n, nrhs))
        for lower in (True, False):
            tri = np.tril(full) if lower else np.triu(full)
            want, want_rc = orc.linsolve_triangular(full, b, lower)
            x, rc = prov.linsolve(prov.upload(full), prov.upload(b), lower=lower, upper=not lower, need_rcond=True)
            got = prov.download(x)
            scale = np.abs(np.linalg.inv(tri)) @ (np.abs(tri) @ np.abs(want) + np.abs(b))   # componentwise forward-error bound of a triangular solve
            assert np.all(np.abs(got - want) <= 1e-10 * scale), (n, nrhs, lower, np.max(np.abs(got - want) / scale))
            assert rc == want_rc
            # TRANSA: solve with A' -- the hint describes A, so the effective triangle flips
            xt, rct = prov.linsolve(prov.upload(np.asfortranarray(full.T)), prov.upload(b), lower=not lower, upper=lower, transposed=True)
            assert np.all(np.abs(prov.download(xt) - want) <= 1e-10 * scale) and rct == want_rc
    # singular / rcond enforcement: the host's message (linsolve.rs:783-786, 1000-1009)
    t = np.diag([1.0, 0.0, 3.0]) + np.tril(np.ones((3, 3)), -1)
    with pytest.raises(ProviderError, match="singular to working precision"):
        prov.linsolve(prov.upload(t), prov.upload(np.ones((3, 1))), lower=True)
    t = np.diag([1.0, 1e-6, 3.0])
    with pytest.raises(ProviderError, match="singular to working precision"):
        prov.linsolve(prov.upload(t), prov.upload(np.ones((3, 1))), upper=True, rcond=1e-3)
    x, rc = prov.linsolve(prov.upload(t), prov.upload(np.ones((3, 1))), upper=True, rcond=1e-9)
    assert abs(rc - 1e-6 / 3.0) < 1e-18
    with pytest.raises(ProviderError, match="square"):
        prov.linsolve(prov.upload(np.ones((3, 2))), prov.upload(np.ones((3, 1))), lower=True)
    with pytest.raises(ProviderError, match="dimensions must agree"):
        prov.linsolve(prov.upload(np.eye(3)), prov.upload(np.ones((2, 1))), lower=True)
    # unstructured / SYM / POSDEF systems: the LU of mldivide when no rcond is requested; otherwise "not supported" -> host fallback
    n = 200
    a = rng.uniform(-1, 1, (n, n)) + np.eye(n) * 8
    b = rng.uniform(-1, 1, (n, 5))
    for kw in ({}, {"transposed": True}, {"symmetric": True}, {"posdef": True}, {"lower": True, "upper": True}):
        aa = a @ a.T if (kw.get("posdef") or kw.get("symmetric")) else a
        x, rc = prov.linsolve(prov.upload(aa), prov.upload(b), **kw)
        eff = aa.T if kw.get("transposed") else aa
        got = prov.download(x)
        assert np.isnan(rc)
        assert np.max(np.abs(eff @ got - b)) <= 1e-10 * (np.max(np.abs(eff)) * np.max(np.abs(got)) * n)
    for kw in ({"need_rcond": True}, {"rcond": 1e-3}, {"rectangular": True}):
        with pytest.raises(ProviderError) as ei:
            prov.linsolve(prov.upload(a), prov.upload(b), **kw)
        assert ei.value.status == 2   # RM_UNSUPPORTED
    assert prov.telemetry_snapshot().linsolve.count > 0


def test_mldivide_kat_and_scalar(prov):
    from svd_solve import mldivide as oracle_mldivide

    a = np.array([1.0, 3.0, 2.0, 4.0]).reshape((2, 2), order="F")  # mldivide.rs:667-668 solves_square_system
    b = np.array([[5.0], [6.0]])
    x = prov.download(prov.mldivide(prov.upload(a), prov.upload(b)))
    assert x.shape == (2, 1) and np.linalg.norm(a @ x - b) < 1e-12
    assert np.allclose(x, oracle_mldivide(a, b), rtol=1e-12, atol=1e-12)
    s = prov.download(prov.mldivide(prov.upload(np.array([[4.0]])), prov.upload(np.array([[2.0, 6.0], [8.0, 1.0]]))))
    assert np.array_equal(s, np.array([[2.0, 6.0], [8.0, 1.0]]) * 0.25)


@pytest.mark.parametrize("n,nrhs", [(1, 1), (3, 2), (64, 1), (65, 7), (130, 33), (257, 5), (1000, 64), (2048, 3)])
def test_mldivide_vs_svd_oracle(prov, n, nrhs):
    from svd_solve import mldivide as oracle_mldivide

    rng = np.random.default_rng(n + nrhs)
    a = rng.uniform(-1, 1, (n, n)) + np.eye(n) * 2.0
    b = rng.uniform(-1, 1, (n, nrhs))
    x = prov.download(prov.mldivide(prov.upload(a), prov.upload(b)))
    want = oracle_mldivide(a, b)
    cond = np.linalg.cond(a)
    assert x.shape == (n, nrhs)
    assert np.linalg.norm(a @ x - b) <= 1e-12 * n * (np.linalg.norm(a) * np.linalg.norm(x) + np.linalg.norm(b))   # backward stable
    assert np.linalg.norm(x - want) <= 1e-13 * cond * n * np.linalg.norm(want) + 1e-300                           # agrees with the SVD solve


def test_mldivide_pivoting_and_fallbacks(prov):
    rng = np.random.default_rng(99)
    n = 200
    a = rng.uniform(-1, 1, (n, n))
    a[np.arange(n), np.arange(n)] = 0.0           # zero diagonal: unusable without row interchanges
    b = rng.uniform(-1, 1, (n, 4))
    x = prov.download(prov.mldivide(prov.upload(a), prov.upload(b)))
    assert np.linalg.norm(a @ x - b) <= 1e-10 * np.linalg.norm(b)
    perm = np.eye(n)[rng.permutation(n)]          # a permutation matrix: every pivot needs a swap, solution is exact
    xp = prov.download(prov.mldivide(prov.upload(perm), prov.upload(b)))
    assert np.array_equal(xp, perm.T @ b)
    with pytest.raises(ProviderError, match="not supported by provider"):   # singular -> host SVD fallback
        prov.mldivide(prov.upload(np.ones((50, 50))), prov.upload(np.ones((50, 1))))
    with pytest.raises(ProviderError, match="not supported by provider"):   # least squares -> host fallback
        prov.mldivide(prov.upload(np.ones((6, 3))), prov.upload(np.ones((6, 1))))
    with pytest.raises(ProviderError, match="same number of rows"):
        prov.mldivide(prov.upload(np.eye(4)), prov.upload(np.ones((5, 1))))
    t = prov.telemetry_snapshot()
    assert t.mldivide.count >= 3


# ---------------------------------------------------------------------------------------------------------------
# a10 / a11: RNG + Monte-Carlo evolution
# ---------------------------------------------------------------------------------------------------------------
def test_random_uniform_stream_is_bit_exact(prov, orc):
    prov.set_rng_state(0)  # 0 -> default seed (simple_provider.rs:3627-3640)
    s = orc.default_seed()
    assert prov.get_rng_state() == s
    for n in (1, 7, 1000, 100003):
        got = prov.download(prov.random_uniform((n, 1)))[:, 0]
        want, s = orc.generate_uniform(s, n)
        assert np.array_equal(got, want)
        assert prov.get_rng_state() == s  # provider state advanced exactly like the host's


def test_random_normal_matches_host_stream(prov, orc):
    prov.set_rng_state(12345)
    s = 12345
    for n in (2, 5, 4096, 100001):
        got = prov.download(prov.random_normal((n, 1)))[:, 0]
        want, s = orc.generate_normal(s, n)
        close(got, want, rtol=1e-10, atol=1e-13)
        assert prov.get_rng_state() == s


def test_stochastic_evolution_zero_scale_kat(prov):
    k = KATS["stochastic_evolution"][0]
    h = prov.upload(np.array(k["state"]).reshape(2, 1))
    got = prov.download(prov.stochastic_evolution(h, k["drift"], k["scale"], k["steps"]))[:, 0]
    want = np.array(k["state"]) * math.exp(k["drift"] * k["steps"])
    assert np.all(np.abs(got - want) < 1e-12)


@pytest.mark.parametrize("n,steps", [(1, 1), (2, 3), (5, 4), (1000, 16), (4097, 8)])
def test_stochastic_evolution_vs_host(prov, orc, n, steps):
    drift, scale = (0.05 - 0.5 * 0.2 ** 2) / 252.0, 0.2 * math.sqrt(1.0 / 252.0)
    s0 = np.full((n, 1), 100.0)
    prov.set_rng_state(777)
    got = prov.download(prov.stochastic_evolution(prov.upload(s0), drift, scale, steps))
    want, new_state = orc.stochastic_evolution(777, s0, drift, scale, steps)
    close(got, want, rtol=1e-10, atol=0)
    assert prov.get_rng_state() == new_state


def test_stochastic_evolution_forms_and_extremes(prov, orc, monkeypatch):
    """The kernel sums the log-returns and exponentiates once when no partial product can leave the f64 range, and multiplies step
    by step (the host's form) otherwise: both forms, the lean and the library math, 256 steps, against the host oracle at 1e-10; a
    drift / scale that overflows must reproduce the host's Inf / 0 pattern."""
    drift, scale = (0.05 - 0.5 * 0.2 ** 2) / 252.0, 0.2 * math.sqrt(1.0 / 252.0)
    s0 = np.linspace(50.0, 150.0, 3001).reshape(-1, 1)
    want, _ = orc.stochastic_evolution(4242, s0, drift, scale, 256)
    for env in ({}, {"RUNMAT_B200_MC_STEPWISE": "1"}, {"RUNMAT_B200_MC_LIBM": "1"}):
        for k in ("RUNMAT_B200_MC_STEPWISE", "RUNMAT_B200_MC_LIBM"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        prov.set_rng_state(4242)
        got = prov.download(prov.stochastic_evolution(prov.upload(s0), drift, scale, 256))
        close(got, want, rtol=1e-10, atol=0)
    for k in ("RUNMAT_B200_MC_STEPWISE", "RUNMAT_B200_MC_LIBM"):
        monkeypatch.delenv(k, raising=False)
    # products that overflow / underflow on the way: span = steps * (|drift| + 8.6 |scale|) >= 600 selects the per-step form
    for d_, sc_, steps in ((40.0, 1.0, 20), (-40.0, 1.0, 20), (0.0, 60.0, 30)):
        s1 = np.array([[1.0], [2.0], [1e-300], [-3.0]])
        prov.set_rng_state(99)
        got = prov.download(prov.stochastic_evolution(prov.upload(s1), d_, sc_, steps))
        want1, _ = orc.stochastic_evolution(99, s1, d_, sc_, steps)
        assert np.array_equal(np.isinf(got), np.isinf(want1)) and np.array_equal(got == 0, want1 == 0) and np.array_equal(np.isnan(got), np.isnan(want1))
        fin = np.isfinite(want1) & (want1 != 0)
        assert np.all(np.abs(got[fin] - want1[fin]) <= 1e-9 * np.abs(want1[fin]))


def test_stochastic_evolution_sharding_is_invariant(prov, orc):
    """Union of shards == single-GPU run, bit for bit."""
    n, steps = 10001, 12
    drift, scale = 1e-4, 0.0126
    s0 = np.random.default_rng(17).uniform(90, 110, (n, 1))
    prov.set_rng_state(42)
    whole = prov.download(prov.stochastic_evolution(prov.upload(s0), drift, scale, steps))
    end_state = prov.get_rng_state()
    for cuts in ([0, 5000, n], [0, 3333, 3334, 7001, n], [0, 1, n]):
        parts = []
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            prov.set_rng_state(42)
            part = prov.stochastic_evolution_sharded(prov.upload(s0[lo:hi]), drift, scale, steps, lo, n)
            parts.append(prov.download(part))
            assert prov.get_rng_state() == end_state
        assert np.array_equal(np.vstack(parts), whole)
    want, _ = orc.stochastic_evolution(42, s0, drift, scale, steps)
    close(whole, want, rtol=1e-10, atol=0)


def test_payoff_partial_sum(prov):
    s = np.random.default_rng(18).uniform(50, 150, (100003, 1))
    got = prov.download(prov.payoff_partial_sum(prov.upload(s), 100.0))[0, 0]
    want = np.maximum(s - 100.0, 0.0).sum()
    assert abs(got - want) <= 1e-12 * want


def test_monte_carlo_full_size_properties(prov):
    """BASELINE config[4] scale on one GPU: 1e8 paths x 256 steps. Zero-scale closed form at full size, and the
    log-return moments of the stochastic run (mean = drift*T, var = scale^2*T) to Monte-Carlo accuracy."""
    M, T = 100_000_000, 256
    h = prov.fill((M, 1), 100.0)
    out = prov.stochastic_evolution(h, 1e-3, 0.0, T)
    tot = prov.download(prov.reduce_sum(out))[0, 0]
    assert abs(tot / M - 100.0 * math.exp(1e-3 * T)) < 1e-9
    prov.free(out)
    drift, scale = (0.05 - 0.5 * 0.2 ** 2) / 252.0, 0.2 * math.sqrt(1.0 / 252.0)
    prov.set_rng_state(0)
    out = prov.stochastic_evolution(h, drift, scale, T)
    lg = prov.unary("log", prov.scalar_div(out, 100.0))
    mean = prov.download(prov.reduce_mean(lg))[0, 0]
    ex2 = prov.download(prov.reduce_mean(prov.elem_mul(lg, lg)))[0, 0]
    var = ex2 - mean * mean
    assert abs(mean - drift * T) < 5 * scale * math.sqrt(T) / math.sqrt(M)
    assert abs(var / (scale * scale * T) - 1.0) < 1e-3


def test_monte_carlo_full_size_sampled_vs_oracle(prov, orc):
    """BASELINE config[4] at full size (1e8 paths x 256 steps) against the oracle on a SAMPLE of paths: the oracle replays the
    sequential host loop for 1000 random paths (plus the first/last pairs) by jumping the host LCG with advance_state."""
    M, T = 100_000_000, 256
    drift, scale = (0.05 - 0.5 * 0.2 ** 2) / 252.0, 0.2 * math.sqrt(1.0 / 252.0)
    h = prov.fill((M, 1), 100.0)
    prov.set_rng_state(0)                     # rng(0) -> the default seed (simple_provider.rs:3627-3640), like the host
    state0 = prov.get_rng_state()
    assert state0 == orc.default_seed()
    out = prov.stochastic_evolution(h, drift, scale, T)
    paths = np.unique(np.concatenate([np.random.default_rng(99).integers(0, M, 1000), [0, 1, 2, 3, M - 2, M - 1, M // 2, M // 2 + 1]])).astype(np.uint32)
    got = prov.download(prov.gather_linear(out, paths, (len(paths), 1)))[:, 0]
    want = orc.stochastic_evolution_sampled(state0, 100.0, M, drift, scale, T, paths)
    assert np.all(np.abs(got - want) <= 1e-10 * np.abs(want)), f"worst {np.max(np.abs(got - want) / np.abs(want))}"
    assert prov.get_rng_state() == orc.advance_state(state0, T * M)   # the provider's RNG advanced exactly like the host's
    prov.free(out)
    prov.free(h)


def _lcg_ops(idx_id, c_mul, c_add, c_mod, base):
    """mod(1664525 .* idx + 1013904223, 2^32) as the planner's op list (Mul, Add, builtin mod)."""
    return [ft.FusionOp("primitive", "ElemMul", [c_mul, idx_id], base), ft.FusionOp("primitive", "Add", [base, c_add], base + 1),
            ft.FusionOp("builtin", "mod", [base + 1, c_mod], base + 2)], base + 2


def test_benchmark_lcg_chains_through_fused_elementwise(prov, orc):
    """The benchmark scripts' own deterministic generators, driven through fused_elementwise (VERDICT r1 missing #3).
    monte-carlo-analysis/runmat_lcg.m:33-45: products 1664525*idx exceed 2^53, so ONE contracted multiply-add changes the
    residue: the states must be BIT-identical to the host's separately rounded multiply, add and mod (-fmad=false).
    4k-image-processing/runmat_lcg.m:59-79: the image field (all integers < 2^53, exact)."""
    two32 = 4294967296.0
    consts = [1664525.0, 1013904223.0, two32]
    hc = [prov.upload(np.array([[c]])) for c in consts]

    def mc_states(M, t, lo, hi, seed=0.0):
        rid = np.arange(lo, hi, dtype=np.float64).reshape(-1, 1)
        salt = float(t) * (2.0 * M)
        idx1, idx2 = rid + salt + seed, rid + salt + float(M) + seed         # script lines 38-39, evaluated left to right
        outs = []
        for idx in (idx1, idx2):
            ops, res = _lcg_ops(0, 1, 2, 3, 10)
            sh = ft.elementwise_wgsl([0, 1, 2, 3], ops, [res])
            got = prov.download(prov.fused_elementwise(sh, [prov.upload(idx)] + hc, idx.shape, idx.size))
            want = orc.elem_binary("mod", 1664525.0 * idx + 1013904223.0, np.array([[two32]]))
            assert np.array_equal(got, want), f"LCG state differs at M={M} t={t}"
            assert np.all(got == np.floor(got)) and got.min() >= 0 and got.max() < two32
            outs.append(got)
        return idx1, idx2, outs

    # reduced M, several steps; then one full-size shard (rank 7 of 8 of M = 1e8) at the last step, where idx ~ 5.1e10
    for t in (0, 1, 7):
        mc_states(100_003, t, 0, 100_003, seed=3.0)
    M = 100_000_000
    idx1, idx2, (s1, s2) = mc_states(M, 255, 7 * M // 8, M)
    assert float(idx2.max()) * 1664525.0 > 2.0 ** 53                           # the FMA hazard is really exercised
    # the rest of the step in f64: u1 = max(state1/2^32, 2^-32); u2 = state2/2^32; z = sqrt(-2 log u1) .* cos(2 pi u2)
    S1, S2, C32, CMIN, CM2, C2PI = 0, 1, 2, 3, 4, 5
    ops = [ft.FusionOp("primitive", "ElemDiv", [S1, C32], 10), ft.FusionOp("builtin", "max", [10, CMIN], 11), ft.FusionOp("primitive", "ElemDiv", [S2, C32], 12),
           ft.FusionOp("builtin", "log", [11], 13), ft.FusionOp("primitive", "ElemMul", [CM2, 13], 14), ft.FusionOp("builtin", "sqrt", [14], 15),
           ft.FusionOp("primitive", "ElemMul", [C2PI, 12], 16), ft.FusionOp("builtin", "cos", [16], 17), ft.FusionOp("primitive", "ElemMul", [15, 17], 18)]
    sh = ft.elementwise_wgsl([S1, S2, C32, CMIN, CM2, C2PI], ops, [18])
    n = 1 << 20
    ins = [prov.upload(s1[:n]), prov.upload(s2[:n])] + [prov.upload(np.array([[c]])) for c in (two32, 1.0 / two32, -2.0, 2.0 * math.pi)]
    z = prov.download(prov.fused_elementwise(sh, ins, (n, 1), n))
    u1 = np.maximum(s1[:n] / two32, 1.0 / two32)
    r = orc.unary("sqrt", -2.0 * orc.unary("log", u1))
    want = r * orc.unary("cos", 2.0 * math.pi * (s2[:n] / two32))
    close(z, want, rtol=1e-10, atol=1e-13)

    # image field: idx = batch_offset + y*W + x with y, x broadcast ([1,H,1] and [1,1,W]); pixel = single(state)/single(2^32)
    for (Bt, H, W, b) in ((3, 37, 53, 2), (64, 2160, 3840, 63)):
        y = np.arange(H, dtype=np.float64).reshape(1, H, 1)
        x = np.arange(W, dtype=np.float64).reshape(1, 1, W)
        Y, Wc, X, OFF, CM, CA, CMOD = 0, 1, 2, 3, 4, 5, 6
        ops = [ft.FusionOp("primitive", "ElemMul", [Y, Wc], 10), ft.FusionOp("primitive", "Add", [OFF, 10], 11), ft.FusionOp("primitive", "Add", [11, X], 12)]
        lcg, res = _lcg_ops(12, CM, CA, CMOD, 13)
        ops += lcg + [ft.FusionOp("primitive", "ElemDiv", [res, CMOD], 20)]
        sh = ft.elementwise_wgsl([Y, Wc, X, OFF, CM, CA, CMOD], ops, [20])
        ins = [prov.upload(y), prov.upload(np.array([[float(W)]])), prov.upload(x), prov.upload(np.array([[float(b * H * W)]]))] + hc
        got = prov.download(prov.fused_elementwise(sh, ins, (1, H, W), H * W)).astype(np.float32)   # exact scaling: round-then-scale == scale-then-round
        want = orc.image_lcg_fill(Bt, H, W, b0=b, bcount=1)
        assert np.array_equal(got, want)


# ---------------------------------------------------------------------------------------------------------------
# a12 / a13: image normalise + imfilter
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,H,W", [(1, 4, 4), (3, 5, 7), (4, 64, 48), (16, 33, 17), (5, 2, 3), (64, 16, 16)])
def test_image_normalize_f64_vs_host(prov, orc, B, H, W):
    x = np.random.default_rng(B * 100 + H).uniform(0, 1, (B, H, W))
    h = prov.upload(x)
    for kw in [dict(gain=1.0123, bias=-0.02, gamma=1.8, clamp_zero=True), dict(gain=None, bias=None, gamma=None, clamp_zero=False),
               dict(gain=2.0, bias=None, gamma=None, clamp_zero=True)]:
        d = ImageNormalizeDescriptor(B, H, W, 1e-6, **kw)
        got = prov.download(prov.image_normalize(h, d))
        want = orc.image_normalize(x, 1e-6, **kw)
        # statistics are accumulated in a different order (and as shifted moments): 1e-10 relative, the pow amplifies slightly
        close(got, want, rtol=1e-9, atol=1e-12)
    with pytest.raises(ProviderError, match="descriptor dims"):
        prov.image_normalize(h, ImageNormalizeDescriptor(B + 1, H, W, 1e-6))
    with pytest.raises(ProviderError, match="epsilon"):
        prov.image_normalize(h, ImageNormalizeDescriptor(B, H, W, -1.0))


def test_image_normalize_constant_image_and_f32(prov, prov32, orc):
    x = np.full((2, 8, 8), 0.25)
    d = ImageNormalizeDescriptor(2, 8, 8, 0.0, clamp_zero=False)
    assert np.array_equal(prov.download(prov.image_normalize(prov.upload(x), d)), np.zeros((2, 8, 8)))  # sigma == 0 -> inv_sigma = 0
    img = orc.image_lcg_fill(4, 120, 160)  # benchmarks/4k-image-processing/runmat_lcg.m:59-79 field, small
    h = prov32.upload(img)
    d = ImageNormalizeDescriptor(4, 120, 160, 1e-6, gain=1.0123, bias=-0.02, gamma=1.8)
    got = prov32.download(prov32.image_normalize(h, d), dtype=np.float32)
    want = orc.image_normalize(img.astype(np.float64), 1e-6, gain=1.0123, bias=-0.02, gamma=1.8)
    assert np.all(np.abs(got - want) <= 5e-4 * np.maximum(np.abs(want), 1.0))  # the reference's f32 tolerance (matmul_small_k.rs:207 style)
    assert np.all(np.abs(got - want) <= 2e-5 * np.maximum(np.abs(want), 1.0))


def test_image_normalize_nan_pixels_and_negative_pow_base(prov, prov32, orc):
    """Rust's value.max(0.0) maps NaN to 0 (simple_provider.rs:7982): a NaN pixel poisons its image's statistics, the clamp then
    zeroes that image and leaves the others untouched. And gamma without the clamp sees negative bases: (-x)^2 = x^2 on both the
    fixed-lane and the generic kernel (exp2(b*log2(a)) returned NaN there)."""
    rng = np.random.default_rng(31)
    for B in (4, 3):                                     # 4: fixed-lane fast path, 3: generic kernel
        x = rng.uniform(0, 1, (B, 16, 24))
        x[1, 5, 7] = np.nan
        for clamp in (True, False):
            d = ImageNormalizeDescriptor(B, 16, 24, 1e-6, gain=1.5, bias=0.1, clamp_zero=clamp)
            got = prov.download(prov.image_normalize(prov.upload(x), d))
            want = orc.image_normalize(x, 1e-6, gain=1.5, bias=0.1, gamma=None, clamp_zero=clamp)
            assert np.array_equal(np.isnan(got), np.isnan(want))
            ok = ~np.isnan(want)
            assert np.all(np.abs(got[ok] - want[ok]) <= 1e-9 * np.abs(want[ok]) + 1e-12)
        x32 = rng.uniform(0, 1, (B, 16, 24)).astype(np.float32)
        d = ImageNormalizeDescriptor(B, 16, 24, 1e-6, gamma=2.0, clamp_zero=False)
        got = prov32.download(prov32.image_normalize(prov32.upload(x32), d), dtype=np.float32)
        want = orc.image_normalize(x32.astype(np.float64), 1e-6, gamma=2.0, clamp_zero=False)
        assert not np.any(np.isnan(got)) and np.min(want) >= 0 and np.any((x32 - x32.mean(axis=(1, 2), keepdims=True)) < 0)
        assert np.all(np.abs(got - want) <= 2e-5 * np.maximum(np.abs(want), 1.0))
        d = ImageNormalizeDescriptor(B, 16, 24, 1e-6, gamma=3.0, clamp_zero=False)      # odd power keeps the sign
        got = prov32.download(prov32.image_normalize(prov32.upload(x32), d), dtype=np.float32)
        want = orc.image_normalize(x32.astype(np.float64), 1e-6, gamma=3.0, clamp_zero=False)
        assert np.all(np.abs(got - want) <= 2e-5 * np.maximum(np.abs(want), 1.0)) and np.any(got < 0)


def test_image_normalize_4k_batch8_f32(prov32, orc):
    """BASELINE config[3] shape (per GPU at 8-way sharding of B=64): [8,2160,3840] f32, checked against the f64 oracle."""
    B, H, W = 8, 2160, 3840
    img = orc.image_lcg_fill(B, H, W)
    h = prov32.upload(img)
    d = ImageNormalizeDescriptor(B, H, W, 1e-6, gain=1.0123, bias=-0.02, gamma=1.8)
    out = prov32.image_normalize(h, d)
    got = prov32.download(out, dtype=np.float32)
    want = orc.image_normalize(img.astype(np.float64), 1e-6, gain=1.0123, bias=-0.02, gamma=1.8)
    assert np.all(np.abs(got - want) <= 2e-5 * np.maximum(np.abs(want), 1.0))
    mse_want = float(np.mean((want - img) ** 2))
    X, Y, T0, T1 = 0, 1, 10, 11
    ops = [ft.FusionOp("primitive", "Sub", [X, Y], T0), ft.FusionOp("primitive", "ElemMul", [T0, T0], T1)]
    sh = ft.reduction_wgsl([X, Y], ops, T1, axis=0, mean=True, scalar_ty="f32")
    mse = prov32.download(prov32.fused_reduction(sh, [out, h], (1, 1), B * H * W, 1, flavor="mean"))[0, 0]
    assert abs(mse - mse_want) <= 1e-5 * mse_want


def test_imfilter_kats_and_modes(prov, orc):
    for k in KATS["imfilter"]:
        o = k["opts"]
        got = prov.download(prov.imfilter(prov.upload(arr(k["img"])), prov.upload(arr(k["ker"])), padding=o.get("padding", "constant"),
                                          shape=o.get("shape", "same"), mode=o.get("mode", "corr")))
        assert_same(got, arr(k["out"]))
    rng = np.random.default_rng(19)
    img = rng.uniform(0, 1, (37, 29, 3))
    g = np.exp(-((np.arange(5) - 2)[:, None] ** 2 + (np.arange(5) - 2)[None, :] ** 2) / 2.0)
    ker = g / g.sum()
    hi, hk = prov.upload(img), prov.upload(ker)
    for padding in ("constant", "replicate", "symmetric", "circular"):
        for shape in ("same", "full", "valid"):
            for mode in ("corr", "conv"):
                got = prov.download(prov.imfilter(hi, hk, padding=padding, constant_value=0.5, shape=shape, mode=mode))
                want = orc.imfilter(img, ker, padding=padding, cval=0.5, shape=shape, mode=mode)
                assert_same(got, want)  # only mul/add in the host's order, no FMA: bit-exact
    # even-sized and rectangular kernels, kernels larger than the image, and a 3-D kernel (generic, non-tiled path)
    for kshape in [(2, 2), (4, 3), (1, 7), (15, 15), (3, 3, 2)]:
        kk = rng.uniform(-1, 1, kshape)
        for padding in ("constant", "symmetric", "circular", "replicate"):
            for shape in ("same", "full", "valid"):
                got = prov.download(prov.imfilter(hi, prov.upload(kk), padding=padding, shape=shape))
                assert_same(got, orc.imfilter(img, kk, padding=padding, shape=shape))
    small = rng.uniform(0, 1, (3, 4))
    got = prov.download(prov.imfilter(prov.upload(small), prov.upload(np.ones((7, 9))), padding="symmetric"))
    assert_same(got, orc.imfilter(small, np.ones((7, 9)), padding="symmetric"))
    big = rng.uniform(0, 1, (300, 200))
    got = prov.download(prov.imfilter(prov.upload(big), hk, padding="replicate"))
    assert_same(got, orc.imfilter(big, ker, padding="replicate"))


def test_imfilter_register_blocked_kernels(prov, prov32, orc):
    """3x3 / 5x5 / 7x7 take the register-blocked kernel: many tiles, ragged edges, every padding/shape/mode, f64 and f32,
    and the result must equal the generic kernel's bit for bit (same per-output tap order)."""
    rng = np.random.default_rng(190)
    img = rng.uniform(-1, 1, (150, 101, 2))
    hi = prov.upload(img)
    for K in (3, 5, 7):
        kk = rng.uniform(-1, 1, (K, K))
        hk = prov.upload(kk)
        for padding in ("constant", "replicate", "symmetric", "circular"):
            for shape in ("same", "full", "valid"):
                for mode in ("corr", "conv"):
                    got = prov.download(prov.imfilter(hi, hk, padding=padding, constant_value=-0.25, shape=shape, mode=mode))
                    assert_same(got, orc.imfilter(img, kk, padding=padding, cval=-0.25, shape=shape, mode=mode))
    # f32 storage: against the f32 oracle (host tap order, a rounding after every multiply and every add), bit for bit
    img32 = rng.uniform(0, 1, (200, 67, 3)).astype(np.float32)
    hi32 = prov32.upload(img32)
    for K in (3, 5, 7):
        k32 = rng.uniform(-1, 1, (K, K)).astype(np.float32)
        for padding in ("constant", "replicate", "symmetric", "circular"):
            got = prov32.download(prov32.imfilter(hi32, prov32.upload(k32), padding=padding, constant_value=0.5), np.float32)
            assert np.array_equal(got, orc.imfilter(img32, k32, padding=padding, cval=0.5, f32=True)), (K, padding)


def test_imfilter_4k_rgb_full_size_vs_oracle(prov32, orc):
    """BASELINE configs[3] frame at full size: 2160x3840x3 f32, 5x5 Gaussian (fspecial sigma=1), 'replicate', 'same' — every
    sample against the f32 oracle, bit for bit (VERDICT r1 missing #5: this config was only timed)."""
    H, W = 2160, 3840
    frame = np.random.default_rng(4).random((H, W, 3), dtype=np.float32)
    g = np.exp(-((np.arange(5) - 2)[:, None] ** 2 + (np.arange(5) - 2)[None, :] ** 2) / 2.0)
    ker = (g / g.sum()).astype(np.float32)
    got = prov32.download(prov32.imfilter(prov32.upload(frame), prov32.upload(ker), padding="replicate"), np.float32)
    want = orc.imfilter(frame, ker, padding="replicate", f32=True)
    assert got.shape == want.shape == (H, W, 3)
    assert np.array_equal(got, want)


def test_kernel_launch_log_and_spawn_policy(prov):
    """ProviderTelemetry::kernel_launches is a bounded log (64 events, newest last) with the wgpu provider's kernel names and shape
    keys (backend/wgpu/provider/ops/telemetry.rs:26-146, helpers.rs:36-50); spawn policy = the in-process provider's."""
    prov.reset_telemetry()
    assert prov.kernel_launch_log() == []
    x = np.random.default_rng(1).uniform(-1, 1, (96, 80))
    h = prov.upload(x)
    one = prov.upload(np.array([[1.0]]))
    c = prov.fused_elementwise(ft.sin_mul_add_wgsl(), [h, h, one], (96, 80), 96 * 80)
    s_ = prov.fused_reduction(ft.sum_sin_mul_add_wgsl(), [h, h], (1, 1), 96 * 80, 1)
    m = prov.matmul(h, prov.upload(x.T.copy()))
    log = prov.kernel_launch_log()
    assert [e["kernel"] for e in log] == ["fused_elementwise", "fused_reduction", "matmul"]
    assert log[0]["shape"] == {"len": 96 * 80, "inputs": 3, "rank": 2} and log[0]["precision"] == "f64"
    assert log[1]["shape"] == {"reduce_len": 96 * 80, "slices": 1, "rank": 2}
    assert log[2]["shape"] == {"m": 96, "n": 96, "k": 80}
    for _ in range(70):
        prov.free(prov.fused_reduction(ft.sum_sin_mul_add_wgsl(), [h, h], (1, 1), 96 * 80, 1))
    log = prov.kernel_launch_log()
    assert len(log) == 64 and all(e["kernel"] == "fused_reduction" for e in log)  # oldest dropped
    assert prov.spawn_handle_concurrency() == "SynchronizedMutation"
    for hh in (h, one, c, s_, m):
        prov.free(hh)


def test_imfilter_tma_staged_path(prov32, orc, monkeypatch):
    """Large f32 images take the persistent TMA-staged kernel (cp.async.bulk.tensor ring for interior tiles, index-map fill for
    the border ring): every padding, 3x3 / 5x5 / 7x7, same / full / valid, corr / conv, ragged tile edges, an RGB and a 2-D
    image -- bit for bit against the f32 oracle and against the per-tile kernel; the pipeline error flag must stay clear."""
    rng = np.random.default_rng(717)
    img = rng.uniform(-1, 1, (1028, 1301, 2)).astype(np.float32)  # 17 x 41 x 2 tiles, ragged in both dims
    hi = prov32.upload(img)
    for K in (3, 5, 7):
        k32 = rng.uniform(-1, 1, (K, K)).astype(np.float32)
        hk = prov32.upload(k32)
        for padding, shape, mode in (("replicate", "same", "corr"), ("constant", "same", "corr"), ("symmetric", "same", "conv"), ("circular", "same", "corr"),
                                     ("replicate", "full", "corr"), ("constant", "valid", "conv")):
            got = prov32.download(prov32.imfilter(hi, hk, padding=padding, constant_value=0.5, shape=shape, mode=mode), np.float32)
            want = orc.imfilter(img, k32, padding=padding, cval=0.5, shape=shape, mode=mode, f32=True)
            assert got.shape == want.shape and np.array_equal(got, want), (K, padding, shape, mode)
        monkeypatch.setenv("RUNMAT_B200_IMFILTER_NO_TMA", "1")
        ref = prov32.download(prov32.imfilter(hi, hk, padding="replicate"), np.float32)
        monkeypatch.delenv("RUNMAT_B200_IMFILTER_NO_TMA")
        assert np.array_equal(ref, prov32.download(prov32.imfilter(hi, hk, padding="replicate"), np.float32))
    flat = rng.uniform(0, 1, (2048, 2600)).astype(np.float32)  # 2-D image: one plane
    k5 = rng.uniform(-1, 1, (5, 5)).astype(np.float32)
    got = prov32.download(prov32.imfilter(prov32.upload(flat), prov32.upload(k5), padding="symmetric"), np.float32)
    assert np.array_equal(got, orc.imfilter(flat, k5, padding="symmetric", f32=True))
    odd = rng.uniform(0, 1, (1030, 1301, 2)).astype(np.float32)  # dim 0 not a multiple of 4: TMA strides illegal -> per-tile kernel
    got = prov32.download(prov32.imfilter(prov32.upload(odd), prov32.upload(k5), padding="replicate"), np.float32)
    assert np.array_equal(got, orc.imfilter(odd, k5, padding="replicate", f32=True))
    assert prov32.device_flags()[0] == 0


def test_conv2d_matches_host_order(prov, orc):
    rng = np.random.default_rng(23)
    for sshape, kshape in [((7, 9), (3, 3)), ((40, 33), (5, 4)), ((2, 2), (3, 5)), ((130, 70), (1, 7)), ((64, 64), (2, 2))]:
        sig, ker = rng.uniform(-1, 1, sshape), rng.uniform(-1, 1, kshape)
        sig[sig > 0.8] = 0.0  # the host skips zero signal entries
        for mode in ("full", "same", "valid"):
            got = prov.download(prov.conv2d(prov.upload(sig), prov.upload(ker), mode))
            want = orc.conv2d(sig, ker, mode)
            assert got.shape == want.shape, (sshape, kshape, mode)
            assert_same(got, want)  # same tap order as the host's scatter loop: bit-exact


def test_explained_variance_hook_sequence(prov, orc):
    """FusionKind::ExplainedVariance (fusion.rs:2481-2583) executes as matmul -> reshape -> matmul -> matmul -> diag_extract
    (fusion_exec.rs:731-869); the "transpose" of Q is a metadata-only reshape there, and is here too."""
    rng = np.random.default_rng(77)
    n = 96
    q, g = rng.uniform(-1, 1, (n, n)), rng.uniform(-1, 1, (n, n))
    g = g @ g.T / n
    hq, hg = prov.upload(q), prov.upload(g)
    live0 = prov.live_buffers()
    tmp0 = prov.matmul(hq, hg)
    qt = prov.reshape(hq, (n, n))
    assert qt.buffer_id == hq.buffer_id  # reshape: same buffer, new logical shape (lib.rs:2676-2684)
    tmp = prov.matmul(qt, hg)
    product = prov.matmul(tmp, prov.reshape(hq, (n, n)))
    diag = prov.diag_extract(product, 0)
    assert diag.shape in ((n, 1), (n,))
    want = np.diagonal(orc.matmul(orc.matmul(q, g), q))
    got = prov.download(diag).reshape(-1)
    scale = np.abs(q) @ np.abs(g) @ np.abs(q)
    assert np.all(np.abs(got - want) <= 1e-10 * np.diagonal(scale))
    for h in (tmp0, tmp, product, diag):
        prov.free(h)
    assert prov.live_buffers() == live0


def test_cat_and_mrdivide(prov):
    rng = np.random.default_rng(24)
    a, b, c = rng.uniform(-1, 1, (3, 4)), rng.uniform(-1, 1, (2, 4)), rng.uniform(-1, 1, (3, 5))
    ha, hb, hc = prov.upload(a), prov.upload(b), prov.upload(c)
    assert np.array_equal(prov.download(prov.cat(1, [ha, hb, ha])), np.concatenate([a, b, a], axis=0))
    assert np.array_equal(prov.download(prov.cat(2, [ha, hc])), np.concatenate([a, c], axis=1))
    assert np.array_equal(prov.download(prov.cat(3, [ha, ha])), np.stack([a, a], axis=2))
    t = rng.uniform(-1, 1, (2, 3, 4))
    assert np.array_equal(prov.download(prov.cat(2, [prov.upload(t), prov.upload(t[:, :1, :])])), np.concatenate([t, t[:, :1, :]], axis=1))
    with pytest.raises(ProviderError, match="dimension mismatch"):
        prov.cat(1, [ha, hc])
    n = 150
    A = rng.uniform(-1, 1, (n, n)) + 3 * np.eye(n)
    Bm = rng.uniform(-1, 1, (7, n))
    X = prov.download(prov.mrdivide(prov.upload(Bm), prov.upload(A)))   # X = Bm / A
    assert X.shape == (7, n) and np.linalg.norm(X @ A - Bm) <= 1e-11 * np.linalg.norm(Bm) * n


# ---------------------------------------------------------------------------------------------------------------
# a15: telemetry
# ---------------------------------------------------------------------------------------------------------------
def test_telemetry_counts(prov):
    prov.reset_telemetry()
    x = np.ones((16, 16))
    h = prov.upload(x)
    h1 = prov.upload(np.array([[1.0]]))
    prov.fused_elementwise(ft.sin_mul_add_wgsl(), [h, h, h1], (16, 16), 256)
    prov.fused_reduction(ft.sum_sin_mul_add_wgsl(), [h, h], (1, 1), 256, 1)
    prov.matmul(h, h)
    prov.download(h)
    t = prov.telemetry_snapshot()
    assert t.fused_elementwise.count == 1 and t.fused_reduction.count == 1 and t.matmul.count == 1
    assert t.upload_bytes == 256 * 8 + 8 and t.download_bytes == 256 * 8
    assert t.kernel_launches >= 3 and t.fused_elementwise.total_wall_time_ns > 0
    assert prov.default_reduction_workgroup_size() == 256 and prov.two_pass_threshold() > 0
    with pytest.raises(ProviderError, match="not supported by provider"):
        prov.mldivide(h, h)


def test_comm_p2p_self_exchange(orc):
    """The peer-memory exchange with world == 1 (a rank connected to itself) runs the whole protocol on one GPU: fused publish
    from the reduction kernel's last block, stand-alone publish, lazy combine on the communication stream, 8-bank reuse."""
    from runmat_b200 import B200Provider
    import fusion_text as ft

    with B200Provider(0, device_id=78) as p:
        assert not p.comm_p2p_connected()
        rng = np.random.default_rng(9)
        n = 1 << 20
        a, b = rng.uniform(0, 4 * np.pi, (n, 1)), rng.uniform(-1, 1, (n, 1))
        ha, hb = p.upload(a), p.upload(b)
        with pytest.raises(ProviderError, match="rm_comm_p2p_connect"):
            p.fused_reduction_allreduce(ft.sum_sin_mul_add_wgsl(), [ha, hb], n)
        h = p.comm_p2p_export()
        assert len(h) == 64
        p.comm_p2p_connect([h], 0, 1)
        assert p.comm_p2p_connected() and p.comm_world_size() == 1
        local = p.download(p.fused_reduction(ft.sum_sin_mul_add_wgsl(), [ha, hb], (1, 1), n, 1))[0, 0]
        want = orc.reduce_sum(orc.sin_mul_add(a, b, 1.0))
        assert abs(local - want) <= 1e-10 * n
        # 40 exchanges in flight-order: cycles every bank 5 times; consumption lags by 4 like the bench loop
        pending = []
        for t in range(40):
            pending.append(p.fused_reduction_allreduce(ft.sum_sin_mul_add_wgsl(), [ha, hb], n))
            if len(pending) > 4:
                assert p.download(pending.pop(0))[0, 0] == local     # world 1: the global sum is the local value, bit for bit
        for hg in pending:
            assert p.download(hg)[0, 0] == local
        # scaled flavour + the stand-alone publish path (1-element tensors through rm_comm_allreduce_sum)
        hg = p.fused_reduction_allreduce(ft.sum_sin_mul_add_wgsl(), [ha, hb], n, flavor="custom", custom_scale=0.25)
        assert abs(p.download(hg)[0, 0] - 0.25 * local) <= 1e-12 * abs(local)
        for v in (3.5, -1e300, 0.0):
            hv = p.upload(np.array([[v]]))
            assert p.download(p.comm_allreduce_sum(hv))[0, 0] == v
        p.free(p.comm_allreduce_sum(p.upload(np.array([[1.0]]))))   # freed before use: ordered after the combine
        p.comm_fence()
        assert p.comm_p2p_error() == 0
        # vectors still need the NCCL communicator
        with pytest.raises(ProviderError, match="rm_comm_init"):
            p.comm_allreduce_sum(ha)


def test_comm_single_rank_roundtrip():
    """rm_comm_* with a one-rank communicator: the all-reduce is the identity, results arrive through the lazy ready event,
    and errors are reported (not raised as crashes) before initialisation."""
    from runmat_b200 import B200Provider

    with B200Provider(0, device_id=77) as p:
        x = np.arange(1000, dtype=np.float64).reshape(1000, 1) * 0.5
        hx = p.upload(x)
        with pytest.raises(ProviderError, match="rm_comm_init"):
            p.comm_allreduce_sum(hx)
        assert p.comm_world_size() == 1
        p.comm_init(B200Provider.comm_unique_id(), 0, 1)
        with pytest.raises(ProviderError, match="already"):
            p.comm_init(B200Provider.comm_unique_id(), 0, 1)
        outs = [p.comm_allreduce_sum(hx) for _ in range(3)]
        p.comm_fence()
        for h in outs:
            assert h.shape == (1000, 1)
            assert np.array_equal(p.download(h), x)
        # the result feeds further device work without an explicit fence (ready-event ordering)
        hs = p.reduce_sum(p.comm_allreduce_sum(p.scalar_mul(hx, 2.0)))
        assert p.download(hs)[0, 0] == float(np.sum(x * 2.0))
        p.free(p.comm_allreduce_sum(hx))  # freed before use: the free is ordered after the collective
        p.synchronize()



def test_interleaved_reduction_layout(prov, prov32, orc, monkeypatch):
    """Reductions over the middle dimension of [inner, n, post] with a small power-of-two `inner` run as one flat vector stream
    per post block (RedLayout::Interleaved): inner below / equal / above the vector width, post > 1, sum / mean / prod, NaN
    include and omit, two-input fused programs, f64 and f32 -- against the oracle and against the Strided kernel."""
    rng = np.random.default_rng(616)
    for P, shape in ((prov, (2, 5001, 3)), (prov, (4, 3000, 2)), (prov, (32, 1000, 5)), (prov, (128, 300)), (prov32, (2, 9000, 3)),
                     (prov32, (4, 8000)), (prov32, (64, 700, 2)), (prov32, (256, 200))):
        f32 = P is prov32
        x = rng.uniform(0.5, 1.5, shape)
        if f32:
            x = x.astype(np.float32)
        x64 = x.astype(np.float64)
        h = P.upload(x)
        want = orc.sum_dims(x64, [1])
        bound = np.abs(x64).sum(axis=1, keepdims=True)
        tol = 1e-6 if f32 else 1e-12
        got = P.download(P.reduce_sum_dim(h, 1))
        assert got.shape == want.shape
        assert np.all(np.abs(got - want) <= tol * bound)
        assert np.array_equal(got, P.download(P.reduce_sum_dim(h, 1)))  # deterministic
        gotm = P.download(P.reduce_mean_dim(h, 1))
        assert np.all(np.abs(gotm - want / shape[1]) <= tol * bound / shape[1])
        monkeypatch.setenv("RUNMAT_B200_RED_NO_INTERLEAVED", "1")
        old = P.download(P.reduce_sum_dim(h, 1))
        monkeypatch.delenv("RUNMAT_B200_RED_NO_INTERLEAVED")
        assert np.all(np.abs(got - old) <= tol * bound)
        # NaN poisons exactly one slice
        xn = x.copy()
        idx = (1, 17) + ((shape[2] - 1,) if len(shape) == 3 else ())
        xn[idx] = np.nan
        hn = P.upload(xn)
        gn = P.download(P.reduce_sum_dim(hn, 1))
        nan_at = (1, 0) + ((shape[2] - 1,) if len(shape) == 3 else ())
        assert np.isnan(gn[nan_at]) and np.count_nonzero(np.isnan(gn)) == 1
        for hh in (h, hn):
            P.free(hh)
    # two-input fused program (x .* y summed along rows, omitnan) on a [128 x 700] matrix: inner = 128 slices
    x, y = rng.uniform(-1, 1, (128, 700)), rng.uniform(-1, 1, (128, 700))
    x[5, 9] = np.nan
    hx, hy = prov.upload(x), prov.upload(y)
    ops = [ft.FusionOp("primitive", "ElemMul", [0, 1], 10)]
    for omit in (False, True):
        sh = ft.reduction_wgsl([0, 1], ops, 10, axis=1, omitnan=omit)
        got = prov.download(prov.fused_reduction(sh, [hx, hy], (128, 1), 700, 128))
        close(got, orc.sum_dims(x * y, [1], omit_nan=omit), rtol=1e-12, atol=1e-13)
    # prod over the middle dimension
    z = rng.uniform(0.99, 1.01, (8, 4000))
    got = prov.download(prov.reduce_prod_dim(prov.upload(z), 1)) if hasattr(prov, "reduce_prod_dim") else None
    if got is not None:
        assert np.all(np.abs(got.reshape(-1) / np.prod(z, axis=1) - 1) <= 1e-12)


def test_few_slices_strided_reduction(prov, prov32, orc):
    """Row-direction reductions with a handful of slices and thousands of rows (many row chunks per slice, folded by the last
    CTA with all threads): sums/means/max/min, NaN include and omit, ragged row counts, f64 and f32."""
    rng = np.random.default_rng(515)
    for P, shape, dims in ((prov, (4, 6001), [1]), (prov, (8, 70, 90), [1, 2]), (prov, (16, 4099), [1]), (prov, (64, 5000), [1]),
                           (prov32, (8, 5003), [1]), (prov32, (8, 72, 91), [1, 2]), (prov32, (16, 4100), [1])):
        f32 = P is prov32
        x = rng.uniform(-1, 1, shape)
        if f32:
            x = x.astype(np.float32)
        h = P.upload(x)
        want = x.astype(np.float64).sum(axis=tuple(dims), keepdims=True)
        bound = np.abs(x.astype(np.float64)).sum(axis=tuple(dims), keepdims=True)
        if len(dims) == 1:
            got = P.download(P.reduce_sum_dim(h, dims[0]))
            ref = P.download(P.reduce_sum_dim(h, dims[0]))
            assert got.shape == want.shape
            assert np.all(np.abs(got - want) <= (1e-6 if f32 else 1e-10) * bound)
            assert np.array_equal(got, ref)  # deterministic
            for is_min in (False, True):
                vals = (P.reduce_min_dim if is_min else P.reduce_max_dim)(h, dims[0])
                v = P.download(vals[0] if isinstance(vals, (tuple, list)) else vals)
                assert np.array_equal(v.reshape(-1), (x.min(axis=1) if is_min else x.max(axis=1)).astype(np.float64).reshape(-1))
        else:
            got = P.download(P.reduce_mean_nd(h, dims))
            n = np.prod([shape[d] for d in dims])
            assert got.shape == want.shape
            assert np.all(np.abs(got - want / n) <= (1e-6 if f32 else 1e-10) * bound / n)
        # NaN include -> NaN in that slice only; omit (fused program) skips it
        xn = x.copy()
        xn[(3,) + tuple(5 for _ in shape[1:])] = np.nan
        hn = P.upload(xn)
        if len(dims) == 1:
            g = P.download(P.reduce_sum_dim(hn, 1)).reshape(-1)
            assert np.isnan(g[3]) and not np.isnan(np.delete(g, 3)).any()
            X = 0
            sh = ft.reduction_wgsl([X], [], X, axis=1, omitnan=True, scalar_ty="f32" if f32 else "f64")
            go = P.download(P.fused_reduction(sh, [hn], (shape[0], 1), shape[1], shape[0])).reshape(-1)
            wo = np.nansum(xn.astype(np.float64), axis=1)
            assert np.all(np.abs(go - wo) <= (1e-6 if f32 else 1e-10) * bound.reshape(-1))


def test_reference_kats_through_the_abi(prov):
    """The reference's own literal test vectors (tests/golden/reference_kats.json: find.rs, sub2ind.rs, ind2sub.rs, permute.rs,
    repmat.rs, cat.rs, eye.rs, mean.rs, prod.rs, max.rs) through the C ABI. Where a reference test only asserts shapes, the
    expected array comes from oracle/layout_ops.py, which the CPU suite pins to the same literals."""
    import importlib.util
    from pathlib import Path

    spec = importlib.util.spec_from_file_location("layout_ops", Path(__file__).resolve().parent.parent / "oracle" / "layout_ops.py")
    lo = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(lo)

    def col_major_strides(dims):
        st, acc = [], 1
        for d in dims:
            st.append(acc)
            acc *= d
        return st

    for k in KATS["find"]:
        res = prov.find(prov.upload(arr(k["a"])), limit=k["limit"], direction=k["direction"])
        lin, rows, cols, vals = (prov.download(r).reshape(-1) for r in res)
        assert lin.tolist() == [float(v) for v in k["linear"]], k["src"]
        if "rows" in k:
            assert rows.tolist() == [float(v) for v in k["rows"]] and cols.tolist() == [float(v) for v in k["cols"]]
            assert vals.tolist() == [float(v) for v in k["vals"]]
    for k in KATS["sub2ind"]:
        subs = [arr(x) for x in k["subs"]]
        n = max(x.size for x in subs)
        mask = [x.size == 1 and n > 1 for x in subs]
        oshape = next((x.shape for x in subs if x.size == n), (1, 1))
        hs = [prov.upload(x) for x in subs]
        if "error" in k:
            with pytest.raises(ProviderError, match="exceeds"):
                prov.sub2ind(k["dims"], col_major_strides(k["dims"]), hs, mask, n, oshape)
            continue
        got = prov.download(prov.sub2ind(k["dims"], col_major_strides(k["dims"]), hs, mask, n, oshape))
        assert_same(got, arr(k["out"]))
    for k in KATS["ind2sub"]:
        idx = arr(k["idx"])
        outs = prov.ind2sub(k["dims"], col_major_strides(k["dims"]), prov.upload(idx), int(np.prod(k["dims"])), idx.size, idx.shape)
        assert len(outs) == len(k["dims"])
        for h, want in zip(outs, k["out"]):
            assert prov.download(h).reshape(-1, order="F").tolist() == [float(v) for v in want], k["src"]
    for k in KATS["permute"]:
        a = arr(k["a"])
        if len(k["order"]) != a.ndim:
            continue  # order longer than the rank (adds a trailing dim) is resolved by the builtin before it reaches a provider
        got = prov.download(prov.permute(prov.upload(a), [o - 1 for o in k["order"]]))
        assert list(got.shape) == k["out_shape"]
        assert np.array_equal(got, lo.permute(a, k["order"]))
    for k in KATS["repmat"]:
        assert_same(prov.download(prov.repmat(prov.upload(arr(k["a"])), k["reps"])), arr(k["out"]))
    for k in KATS["cat"]:
        assert_same(prov.download(prov.cat(k["dim"], [prov.upload(arr(x)) for x in k["inputs"]])), arr(k["out"]))
    for k in KATS["eye"]:
        assert_same(prov.download(prov.eye((k["rows"], k["cols"]))), arr(k["out"]))
    for k in KATS["mean"]:
        if k["omitnan"] or len(k["dims"]) != 1:
            continue  # omitnan goes through fused_reduction (covered by test_fused_reduction_axes_omitnan_mean)
        assert_same(prov.download(prov.reduce_mean_dim(prov.upload(arr(k["a"])), k["dims"][0])), arr(k["out"]))
    for k in KATS["prod"]:
        a = arr(k["a"])
        if len(k["dims"]) == a.ndim:  # the trait only has the 'all' product (lib.rs:2733)
            assert prov.download(prov.reduce_prod(prov.upload(a)))[0, 0] == arr(k["out"]).reshape(-1)[0]
    for k in KATS["max_dim"]:
        v, i = prov.reduce_max_dim(prov.upload(arr(k["a"])), k["dim"])
        assert prov.download(v).reshape(-1).tolist() == [float(x) for x in k["values"]]
        assert prov.download(i).reshape(-1).tolist() == [float(x) for x in k["indices"]]
