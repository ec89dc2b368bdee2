"""Pins the CPU oracle against the literal known-answer vectors of the reference's own unit tests."""
import math

import numpy as np
import pytest

from kat_util import KATS, arr, assert_same, num


def test_elem_binary_kats(orc):
    for k in KATS["elem_binary"]:
        assert_same(orc.elem_binary(k["op"], arr(k["a"]), arr(k["b"])), arr(k["out"]))


@pytest.mark.parametrize("name", ["mod", "rem"])
def test_mod_rem_kats(orc, name):
    fn = orc.mod if name == "mod" else orc.rem
    for a, b, want in KATS[name]:
        got, want = fn(num(a), num(b)), num(want)
        assert (math.isnan(got) and math.isnan(want)) or got == want, (name, a, b, got, want)
        # the array path agrees with the scalar path
        got2 = orc.elem_binary(name, np.array([[num(a)]]), np.array([[num(b)]]))[0, 0]
        assert (math.isnan(got2) and math.isnan(want)) or got2 == want


def test_mod_negative_zero_normalised(orc):
    assert math.copysign(1.0, orc.mod(-4.0, 2.0)) == 1.0
    assert math.copysign(1.0, orc.rem(-4.0, 2.0)) == 1.0


def test_sum_kats(orc):
    for k in KATS["sum"]:
        assert_same(orc.sum_dims(arr(k["a"]), k["dims"], k["omitnan"]), arr(k["out"]))


def test_matmul_kats(orc):
    for k in KATS["matmul"]:
        assert_same(orc.matmul(arr(k["a"]), arr(k["b"]), naive=True), arr(k["out"]))
        assert_same(orc.matmul(arr(k["a"]), arr(k["b"]), naive=False), arr(k["out"]))


def test_matmul_interchanged_is_bit_identical_to_naive(orc):
    rng = np.random.default_rng(0)
    for m, k, n in [(1, 1, 1), (7, 13, 5), (64, 257, 33), (130, 96, 70)]:
        a, b = rng.uniform(-1, 1, (m, k)), rng.uniform(-1, 1, (k, n))
        assert np.array_equal(orc.matmul(a, b, naive=True), orc.matmul(a, b, naive=False))


def test_matmul_epilogue_kat(orc):
    k = KATS["matmul_epilogue"][0]
    a, b = arr(k["a"]), arr(k["b"])
    base = orc.matmul(a, b, naive=True)
    got, _ = orc.matmul_epilogue(base, alpha=k["alpha"], beta=k["beta"], row_scale=k["row_scale"], col_scale=k["col_scale"])
    want = (base * k["alpha"] + k["beta"]) * np.array(k["row_scale"])[:, None] * np.array(k["col_scale"])[None, :]
    assert_same(got, want, tol=1e-9)  # the reference test's tolerance


def test_matmul_epilogue_order_clamp_pow_diag(orc):
    # simple_provider.rs:7805-7838: alpha/beta -> row -> col -> clamp_min -> clamp_max -> pow -> diag
    c = np.array([[1.0, -2.0], [3.0, 4.0]])
    diag0 = np.zeros(2)
    got, diag = orc.matmul_epilogue(c, alpha=2.0, beta=1.0, row_scale=[2.0, 4.0], row_div=True, clamp_min=0.0, clamp_max=2.0,
                                    pow_exponent=2.0, diag=diag0)
    want = np.minimum(np.maximum((c * 2.0 + 1.0) / np.array([2.0, 4.0])[:, None], 0.0), 2.0) ** 2.0
    assert_same(got, want)
    assert_same(diag, np.array([want[0, 0], want[1, 1]]))


def test_imfilter_kats(orc):
    for k in KATS["imfilter"]:
        o = k["opts"]
        got = orc.imfilter(arr(k["img"]), arr(k["ker"]), padding=o.get("padding", "constant"), shape=o.get("shape", "same"), mode=o.get("mode", "corr"))
        assert_same(got, arr(k["out"]), tol=1e-12)


def test_imfilter_conv_equals_corr_with_flipped_kernel(orc):
    # imfilter.rs:865-894
    img = np.array([1.0, 4, 2, 5, 3, 6]).reshape((3, 2), order="F")
    ker = np.array([1.0, 2, 3, 4]).reshape((2, 2), order="F")
    flipped = np.array([4.0, 3, 2, 1]).reshape((2, 2), order="F")
    assert_same(orc.imfilter(img, ker, mode="conv"), orc.imfilter(img, flipped, mode="corr"))


def test_imfilter_symmetric_padding_reflects(orc):
    img = np.arange(1.0, 7.0).reshape((2, 3), order="F")
    ker = np.ones((3, 3))
    got = orc.imfilter(img, ker, padding="symmetric")
    padded = np.pad(img, 1, mode="reflect")  # period 2*len-2 (imfilter.rs:777-792)
    want = np.array([[padded[i:i + 3, j:j + 3].sum() for j in range(3)] for i in range(2)])
    assert_same(got, want)


def test_stochastic_evolution_zero_scale_kat(orc):
    k = KATS["stochastic_evolution"][0]
    got, _ = orc.stochastic_evolution(orc.default_seed(), np.array(k["state"]).reshape(2, 1), k["drift"], k["scale"], k["steps"])
    want = np.array(k["state"]) * math.exp(k["drift"] * k["steps"])
    assert np.all(np.abs(got.ravel() - want) < 1e-12)


def test_linspace_kat(orc):
    k = KATS["linspace"][0]
    assert_same(orc.linspace(k["start"], k["stop"], k["count"]).ravel(), np.array(k["out"]))
    assert orc.linspace(0.0, 4 * math.pi, 1000)[0, -1] == 4 * math.pi


def test_rng_constants_and_jump_ahead(orc):
    r = KATS["rng"]
    assert orc.default_seed() == r["default_seed"]
    assert orc.mix_seed(0) == r["default_seed"]
    s = r["default_seed"]
    # one LCG step by hand: state*mult+inc mod 2^64, >>11, *2^-53 (random.rs:271-277)
    s1 = (s * r["multiplier"] + r["increment"]) % (1 << 64)
    u, new = orc.generate_uniform(s, 1)
    assert new == s1 and u[0] == (s1 >> r["shift"]) * (1.0 / (1 << 53))
    # advance_state(state, n) == n sequential steps (random.rs:238-257)
    for n in (0, 1, 2, 3, 17, 1000, 12345):
        _, seq = orc.generate_uniform(s, n)
        assert orc.advance_state(s, n) == seq
    u = orc.generate_uniform(s, 4096)[0]
    assert np.all((u >= 0.0) & (u < 1.0))


def test_normals_are_box_muller_pairs(orc):
    s = orc.default_seed()
    z, new = orc.generate_normal(s, 5)  # odd: the last pair still consumes two uniforms
    u, new_u = orc.generate_uniform(s, 6)
    assert new == new_u
    r0 = math.sqrt(-2.0 * math.log(u[0]))
    assert z[0] == r0 * math.cos(2.0 * math.pi * u[1]) and z[1] == r0 * math.sin(2.0 * math.pi * u[1])
    big = orc.generate_normal(s, 200000)[0]
    assert abs(big.mean()) < 0.01 and abs(big.std() - 1.0) < 0.01  # runmat-runtime/tests/rng.rs:21-62 style moments


def test_sampled_evolution_equals_the_sequential_host_loop(orc):
    """The sampled replay used for the full-size Monte-Carlo parity check is the sequential host loop, bit for bit, for odd and
    even lengths (a trailing unpaired element still consumes a whole Box-Muller pair)."""
    drift, scale = 1.2e-4, 0.0126
    for n, steps in ((1, 3), (2, 2), (7, 5), (1000, 9), (1001, 4)):
        s0 = np.full((n, 1), 100.0)
        want, _ = orc.stochastic_evolution(4242, s0, drift, scale, steps)
        got = orc.stochastic_evolution_sampled(4242, 100.0, n, drift, scale, steps, np.arange(n))
        assert np.array_equal(got, want[:, 0])
    # NaN pixels: Rust's value.max(0.0) maps NaN to 0 (simple_provider.rs:7982); the restatement must too
    x = np.random.default_rng(1).uniform(0, 1, (2, 4, 4))
    x[1, 2, 3] = np.nan
    out = orc.image_normalize(x, 1e-6, gain=1.5, bias=0.1, gamma=None, clamp_zero=True)
    assert np.all(out[1] == 0.0) and np.all(np.isfinite(out[0]))     # a NaN pixel poisons its image's statistics -> every pixel clamps to 0
    out = orc.image_normalize(x, 1e-6, gain=1.5, bias=0.1, gamma=None, clamp_zero=False)
    assert np.all(np.isnan(out[1])) and np.all(np.isfinite(out[0]))
    # f32 imfilter oracle: same tap order, f32 arithmetic; equal to the f64 one on dyadic data
    img = np.random.default_rng(2).integers(0, 64, (9, 7, 2)).astype(np.float64) / 64.0
    ker = np.random.default_rng(3).integers(-8, 8, (3, 3)).astype(np.float64) / 8.0
    for padding in ("constant", "replicate", "symmetric", "circular"):
        assert np.array_equal(orc.imfilter(img, ker, padding=padding, f32=True).astype(np.float64), orc.imfilter(img, ker, padding=padding))


def test_linsolve_triangular_kats(orc):
    for k in KATS["linsolve"]:
        o = k["opts"]
        if not (o.get("lower") or o.get("upper")):
            continue
        a = arr(k["a"])
        lower = bool(o.get("lower"))
        if o.get("transposed"):
            a, lower = np.asfortranarray(a.T), not lower          # linsolve.rs:698-705: transpose, then the triangle hint flips
        x, rc = orc.linsolve_triangular(a, arr(k["b"]), lower)
        assert_same(x, arr(k["out"]), tol=1e-7)                     # the reference's approx_eq
        if "rcond" in k:
            assert abs(rc - k["rcond"]) < 1e-7
    with pytest.raises(ValueError):
        orc.linsolve_triangular(np.array([[1.0, 0.0], [2.0, 0.0]]), np.ones((2, 1)), True)
    # the untouched triangle is ignored, like the host loops
    t = np.array([[2.0, 99.0], [1.0, 4.0]])
    x, rc = orc.linsolve_triangular(t, np.array([[2.0], [9.0]]), True)
    assert np.array_equal(x, [[1.0], [2.0]]) and rc == 0.5


def test_unary_and_scalar_tables(orc):
    x = np.array([[-2.5, -0.0, 0.0, 0.5, 2.5, np.nan, np.inf]])
    assert_same(orc.unary("round", x), np.array([[-3.0, -0.0, 0.0, 1.0, 3.0, np.nan, np.inf]]))  # half away from zero
    assert_same(orc.unary("sign", x), np.array([[-1.0, 0.0, 0.0, 1.0, 1.0, np.nan, 1.0]]))
    assert_same(orc.unary("heaviside", x), np.array([[0.0, 0.5, 0.5, 1.0, 1.0, np.nan, 1.0]]))
    assert_same(orc.unary("fix", x), np.array([[-2.0, -0.0, 0.0, 0.0, 2.0, np.nan, np.inf]]))
    assert_same(orc.scalar_op("rsub", np.array([[1.0, 2.0]]), 10.0), np.array([[9.0, 8.0]]))
    assert_same(orc.scalar_op("rdiv", np.array([[2.0, 4.0]]), 8.0), np.array([[4.0, 2.0]]))
    # elementwise max/min: NaN loses unless both NaN (max.rs:2323-2343)
    assert_same(orc.elem_binary("max", np.array([[np.nan, 1.0, np.nan]]), np.array([[2.0, np.nan, np.nan]])), np.array([[2.0, 1.0, np.nan]]))


def test_broadcast_error_and_zero_dims(orc):
    with pytest.raises(ValueError):
        orc.elem_binary("add", np.ones((2, 3)), np.ones((3, 2)))
    assert orc.elem_binary("add", np.ones((0, 3)), np.ones((1, 3))).shape == (0, 3)


def test_image_normalize_matches_test_restatement(orc):
    # runmat-accelerate/tests/image_normalize.rs:7-66 restates the provider code; check against a numpy form
    rng = np.random.default_rng(3)
    x = rng.uniform(0, 1, (3, 5, 7))
    got = orc.image_normalize(x, 1e-6, gain=1.0123, bias=-0.02, gamma=1.8)
    mu = x.mean(axis=(1, 2), keepdims=True)
    sig = np.sqrt(((x - mu) ** 2).mean(axis=(1, 2), keepdims=True) + 1e-6)
    want = np.maximum((x - mu) / sig * 1.0123 - 0.02, 0.0) ** 1.8
    assert_same(got, want, tol=1e-12)


def test_conv2d_reference_kats(orc):
    """Literal expectations of the reference's conv2 unit tests (runmat-runtime/src/builtins/math/signal/conv2.rs:736-978;
    `tensor_from_rows` there takes row-major literals). They pin the reference's kernel orientation and 'same' alignment."""
    rows = lambda r, c, d: np.array(d, dtype=np.float64).reshape(r, c)  # noqa: E731
    a22, ones22 = rows(2, 2, [1, 2, 3, 4]), rows(2, 2, [1, 1, 1, 1])
    assert_same(orc.conv2d(a22, ones22, "full"), rows(3, 3, [1, 3, 2, 4, 10, 6, 3, 7, 4]))                       # :736 conv2_full_basic
    a33 = rows(3, 3, [1, 2, 3, 4, 5, 6, 7, 8, 9])
    assert_same(orc.conv2d(a33, np.ones((3, 3)), "same"), rows(3, 3, [12, 21, 16, 27, 45, 33, 24, 39, 28]))     # :753 conv2_same_matches_reference
    assert_same(orc.conv2d(a33, rows(3, 3, [1, 0, -1, 1, 0, -1, 1, 0, -1]), "same"),
                rows(3, 3, [-7, -4, 7, -15, -6, 15, -13, -4, 13]))                                              # :778 conv2_same_flips_kernel
    assert_same(orc.conv2d(a33, np.ones((3, 3)), "valid"), np.array([[45.0]]))                                  # :803 conv2_valid_returns_expected_sum
    assert_same(orc.conv2d(a33, rows(2, 2, [1, 2, 3, 4]), "same"), rows(3, 3, [4, 11, 18, 18, 37, 47, 36, 67, 77]))  # :948 even kernel alignment
    assert orc.conv2d(np.ones((2, 2)), np.ones((3, 3)), "valid").size == 0


# ---- reductions beyond sum: literals from mean.rs / prod.rs / max.rs ---------------------------------------------------
def _mean_dims(orc, a, dims, omit):
    """mean.rs:1121-1154: sum of the kept elements / their count (omitnan: count of non-NaN; an all-NaN slice gives NaN)."""
    s = orc.sum_dims(a, dims, omit_nan=omit)
    if omit:
        cnt = orc.sum_dims((~np.isnan(a)).astype(np.float64), dims)
        with np.errstate(invalid="ignore", divide="ignore"):
            return np.where(cnt > 0, s / cnt, np.nan)
    n = np.prod([a.shape[d] for d in dims])
    return s / n


def test_mean_kats(orc):
    for k in KATS["mean"]:
        a = arr(k["a"])
        assert_same(_mean_dims(orc, a, k["dims"], k["omitnan"]), arr(k["out"]), tol=1e-12 if k["omitnan"] else 0.0)
        if not k["omitnan"] and len(k["dims"]) == a.ndim:
            assert_same(np.array(orc.reduce_mean(a)), arr(k["out"]).reshape(()))


def test_prod_kats(orc):
    for k in KATS["prod"]:
        a, want = arr(k["a"]), arr(k["out"])
        keep = [d for d in range(a.ndim) if d not in k["dims"]]
        moved = np.transpose(a, keep + list(k["dims"]))  # slices first, reduced dims last; column-major order inside a slice
        got = np.array([orc.reduce_prod(np.asfortranarray(moved[idx]).reshape(-1, order="F")) for idx in np.ndindex(*[a.shape[d] for d in keep])])
        assert_same(got.reshape(want.shape, order="F") if keep else got.reshape(want.shape), want)


def test_max_dim_kats_values_and_first_index(orc):
    for k in KATS["max_dim"]:
        a = arr(k["a"])
        vals, idx = orc.reduce_minmax_dim(a, k["dim"], is_min=False)
        assert_same(vals.reshape(-1), np.array(k["values"], dtype=np.float64))
        assert_same(idx.reshape(-1), np.array(k["indices"], dtype=np.float64))  # 1-based, first occurrence


# ---- indexing / layout class: numpy restatement (oracle/layout_ops.py) pinned to the reference's literals ----------------
import importlib.util as _ilu
from pathlib import Path as _Path

_spec = _ilu.spec_from_file_location("layout_ops", _Path(__file__).resolve().parent.parent / "oracle" / "layout_ops.py")
layout_ops = _ilu.module_from_spec(_spec)
_spec.loader.exec_module(layout_ops)


def test_find_kats():
    for k in KATS["find"]:
        lin, rows, cols, vals = layout_ops.find(arr(k["a"]), limit=k["limit"], direction=k["direction"])
        assert lin.tolist() == [float(v) for v in k["linear"]]
        if "rows" in k:
            assert rows.tolist() == k["rows"] and cols.tolist() == k["cols"] and vals.tolist() == k["vals"]
    lin, *_ = layout_ops.find(np.array([[0.0, np.nan], [-0.0, 1e-300]]))  # NaN and denormals are non-zero, -0.0 is zero
    assert lin.tolist() == [3.0, 4.0]
    assert layout_ops.find(np.zeros((3, 2)))[0].size == 0 and layout_ops.find(np.ones((2, 2)), limit=0)[0].size == 0


def test_sub2ind_ind2sub_kats():
    for k in KATS["sub2ind"]:
        subs = [arr(s) for s in k["subs"]]
        if "error" in k:
            with pytest.raises(ValueError, match=k["error"]):
                layout_ops.sub2ind(k["dims"], subs)
            continue
        assert_same(layout_ops.sub2ind(k["dims"], subs), arr(k["out"]))
    for k in KATS["ind2sub"]:
        outs = layout_ops.ind2sub(k["dims"], arr(k["idx"]))
        assert len(outs) == len(k["dims"])
        for o, w in zip(outs, k["out"]):
            assert o.shape == tuple(k["idx"]["shape"]) and o.reshape(-1, order="F").tolist() == [float(v) for v in w]
    # round trip and error wording (ind2sub.rs:326-352)
    dims = [4, 3, 5]
    idx = np.arange(1, 61, dtype=np.float64).reshape(6, 10, order="F")
    assert_same(layout_ops.sub2ind(dims, layout_ops.ind2sub(dims, idx)), idx)
    for bad, msg in ((0.0, "positive integers"), (2.5, "positive integers"), (61.0, "must not exceed 60")):
        with pytest.raises(ValueError, match=msg):
            layout_ops.ind2sub(dims, np.array([[bad]]))


def test_permute_repmat_cat_eye_kats():
    for k in KATS["permute"]:
        a = arr(k["a"])
        out = layout_ops.permute(a, k["order"])
        assert list(out.shape) == k["out_shape"]
        inv = np.argsort([o - 1 for o in k["order"]])
        assert np.array_equal(np.transpose(out, inv).reshape(-1, order="F")[: a.size], a.reshape(-1, order="F"))
    for k in KATS["repmat"]:
        assert_same(layout_ops.repmat(arr(k["a"]), k["reps"]), arr(k["out"]))
    for k in KATS["cat"]:
        assert_same(layout_ops.cat(k["dim"], [arr(x) for x in k["inputs"]]), arr(k["out"]))
    with pytest.raises(ValueError, match="dimension mismatch"):
        layout_ops.cat(1, [np.zeros((2, 2)), np.zeros((2, 3))])
    for k in KATS["eye"]:
        assert_same(layout_ops.eye(k["rows"], k["cols"]), arr(k["out"]))
