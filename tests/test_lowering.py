"""CPU tests of the host logic: WGSL parse -> CUDA lowering -> NVRTC compile for sm_100a (no device needed)."""
import ctypes as C

import pytest

import fusion_text as ft


@pytest.fixture(scope="module")
def lib(built):
    from runmat_b200 import _capi

    return _capi.lib


def lower_ew(lib, shader, variant=0, mask=0):
    need = C.c_size_t()
    st = lib.rm_debug_lower_elementwise(shader.encode(), variant, mask, None, 0, C.byref(need))
    if st != 0:
        raise RuntimeError(lib.rm_last_error().decode())
    buf = C.create_string_buffer(need.value)
    lib.rm_debug_lower_elementwise(shader.encode(), variant, mask, buf, need.value, None)
    return buf.value.decode()


def lower_red(lib, shader, op=0, layout=-1):
    need, axis, omit = C.c_size_t(), C.c_int(), C.c_int()
    st = lib.rm_debug_lower_reduction(shader.encode(), op, layout, None, 0, C.byref(need), C.byref(axis), C.byref(omit))
    if st != 0:
        raise RuntimeError(lib.rm_last_error().decode())
    buf = C.create_string_buffer(need.value)
    lib.rm_debug_lower_reduction(shader.encode(), op, layout, buf, need.value, None, None, None)
    return buf.value.decode(), axis.value, omit.value


def translate(lib, expr, ty="f64"):
    buf = C.create_string_buffer(4096)
    st = lib.rm_debug_translate_expr(expr.encode(), ty.encode(), buf, 4096)
    if st != 0:
        raise RuntimeError(lib.rm_last_error().decode())
    return buf.value.decode()


def compiles(lib, src):
    sz = C.c_size_t()
    st = lib.rm_debug_compile(src.encode(), b"t.cu", C.byref(sz))
    assert st == 0, lib.rm_last_error().decode()[:2000]
    return sz.value


def test_expression_translation(lib):
    assert translate(lib, "(input0.data[i0] * input1.data[i1])") == "(v0*v1)"
    assert translate(lib, "sin(input0.data[i0])") == "rm_sin(v0)"    # f64: the prelude's lean sin; f32 / RUNMAT_B200_LIBM_TRIG: the library's
    assert translate(lib, "(log(tmp3) * f64(0.4342944819032518))") == "(log(tmp3)*rm_f64(0.4342944819032518))"
    assert translate(lib, "select(f64(0.0), f64(1.0), (tmp0 > f64(0.0)))") == "rm_select(rm_f64(0.0), rm_f64(1.0), (tmp0>rm_f64(0.0)))"
    assert translate(lib, "pow(v, f64(2))") == "pow(v0, rm_f64(2.0))"          # integer literal stays exact
    assert translate(lib, "(v1 + 0.5)", "f32") == "(v1+0.5f)"                    # abstract float takes the scalar type
    assert translate(lib, "abs(min(v, v1))") == "fabs(fmin(v0, v1))"
    assert "rm_isinf(v1)&&rm_isfinite(v0)" in translate(lib, "(isInf(v1) && isFinite(v))")
    for bad in ["frobnicate(v)", "params.len", "v[3]", "3u", "v % v1", "foo"]:
        with pytest.raises(RuntimeError):
            translate(lib, bad)


def test_elementwise_parse_and_compile_all_variants(lib):
    for ty in ("f64", "f32"):
        sh = ft.sin_mul_add_wgsl(ty)
        src = lower_ew(lib, sh, 0, 0b100)
        assert "const T s2 = in2[0];" in src and "const T tmp0 = rm_sin(v0);" in src and "ld.global.nc.L1::no_allocate" in src
        assert compiles(lib, src) > 0
        assert compiles(lib, lower_ew(lib, sh, 1, 0)) > 0


def test_lean_trig_against_libm(lib, tmp_path):
    """The generated kernels' f64 sin / cos (prelude block "rm_trig") compiled for the host and compared with glibc: <= 3 ulp over
    the fast range |x| < 105615 (measured 2.4), fdlibm answers for tiny arguments, the library routine beyond, NaN for Inf / NaN.
    The parity bar against the reference's CPU sin / cos (f64::sin, sin.rs) is 1e-10 relative (tests/test_gpu_parity.py)."""
    import pathlib
    import subprocess

    src = lower_ew(lib, ft.sin_mul_add_wgsl(), 0, 0b100)
    a, b = src.index("// ---- rm_trig begin"), src.index("// ---- rm_trig end")
    (tmp_path / "fused_trig_block.inc").write_text(src[a:b])
    cpp = pathlib.Path(__file__).parent / "golden" / "check_fused_trig.cpp"
    exe = tmp_path / "check_fused_trig"
    subprocess.run(["g++", "-O2", "-ffp-contract=off", f"-I{tmp_path}", "-o", str(exe), str(cpp)], check=True)
    r = subprocess.run([str(exe), "3.0"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    # f32 programs keep the library functions
    assert "#define rm_sin(x) sin(x)" in lower_ew(lib, ft.sin_mul_add_wgsl("f32"), 0, 0b100)


def test_libm_trig_switch_lowers_and_compiles():
    """RUNMAT_B200_LIBM_TRIG=1 (the A/B switch of the lean sin / cos) is read when the source is emitted: in a fresh process the
    generated kernels must define RM_LIBM_TRIG 1, map rm_sin to the library's sin, and still compile for sm_100a."""
    import os
    import subprocess
    import sys

    code = (
        "import sys, ctypes as C\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from runmat_b200 import _capi\n"
        "import fusion_text as ft\n"
        "lib = _capi.lib\n"
        "for red in (0, 1):\n"
        "    need = C.c_size_t()\n"
        "    if red:\n"
        "        lib.rm_debug_lower_reduction(ft.sum_sin_mul_add_wgsl().encode(), 0, 0, None, 0, C.byref(need), None, None)\n"
        "        buf = C.create_string_buffer(need.value)\n"
        "        lib.rm_debug_lower_reduction(ft.sum_sin_mul_add_wgsl().encode(), 0, 0, buf, need.value, None, None, None)\n"
        "    else:\n"
        "        lib.rm_debug_lower_elementwise(ft.sin_mul_add_wgsl().encode(), 0, 4, None, 0, C.byref(need))\n"
        "        buf = C.create_string_buffer(need.value)\n"
        "        lib.rm_debug_lower_elementwise(ft.sin_mul_add_wgsl().encode(), 0, 4, buf, need.value, None)\n"
        "    src = buf.value.decode()\n"
        "    assert '#define RM_LIBM_TRIG 1' in src and 'rm_sin(v0)' in src\n"
        "    assert lib.rm_debug_compile(buf.value, b't.cu', None) == 0, lib.rm_last_error()\n"
        "print('ok')\n"
    ) % (str(__import__("pathlib").Path(__file__).resolve().parent.parent), str(__import__("pathlib").Path(__file__).resolve().parent))
    env = dict(os.environ, RUNMAT_B200_LIBM_TRIG="1", RUNMAT_B200_NO_KCACHE="1")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


def test_multi_output_and_every_builtin_lower(lib):
    names = ["sin", "cos", "tan", "asin", "acos", "atan", "sinh", "cosh", "tanh", "exp", "log", "log2", "sqrt", "abs", "exp2", "floor", "ceil",
             "round", "trunc", "fix", "sign", "heaviside", "isnan", "isinf", "isfinite", "single", "double", "log10", "log1p", "expm1", "asinh",
             "acosh", "atanh", "pow2", "gpuarray"]  # every name of the planner tables (fusion.rs:2874-3026)
    ops, vid = [], 100
    for nme in names:
        ops.append(ft.FusionOp("builtin", nme, [0], vid))
        vid += 1
    for nme in ["mod", "rem", "atan2", "hypot", "max", "min"]:
        ops.append(ft.FusionOp("builtin", nme, [0, 1], vid))
        vid += 1
    for nme in ["Add", "Sub", "ElemMul", "ElemDiv", "ElemPow"]:
        ops.append(ft.FusionOp("primitive", nme, [0, 1], vid))
        vid += 1
    ops.append(ft.FusionOp("primitive", "Neg", [0], vid))
    sh = ft.elementwise_wgsl([0, 1], ops, [vid, 100])
    src = lower_ew(lib, sh, 0, 0)
    assert "T* __restrict__ out0, T* __restrict__ out1" in src
    assert compiles(lib, src) > 0


def test_reduction_parse_axes_omitnan_and_compile(lib):
    ops = [ft.FusionOp("primitive", "ElemMul", [0, 1], 10)]
    for axis in (0, 1):
        for omit in (False, True):
            sh = ft.reduction_wgsl([0, 1], ops, 10, axis=axis, omitnan=omit)
            src, ax, om = lower_red(lib, sh)
            assert ax == axis and om == int(omit)
            assert "(v0*v1)" in src
    for op in range(4):
        for layout in (0, 1, 2):
            for ty in ("f64", "f32"):
                src = lower_red(lib, ft.sum_sin_mul_add_wgsl(ty), op, layout)[0]
                assert compiles(lib, src) > 0
                if layout == 2:  # interleaved layout: per-lane accumulators, xor-shuffle fold, no floating-point atomics
                    assert "acc[VEC]" in src and "__shfl_xor_sync" in src and "atomicAdd(&tickets" in src and "atomicAdd(&sh" not in src


def test_parse_errors(lib):
    with pytest.raises(RuntimeError, match="struct Tensor"):
        lower_ew(lib, "fn main() {}")
    sh = ft.sin_mul_add_wgsl()
    with pytest.raises(RuntimeError, match="no output store"):
        lower_ew(lib, sh.replace("output.data[g]", "// output"))
    with pytest.raises(RuntimeError, match="unsupported scalar type"):
        lower_ew(lib, sh.replace("array<f64>", "array<i32>"))
    with pytest.raises(RuntimeError, match="axis"):
        lower_red(lib, ft.sum_sin_mul_add_wgsl().replace("let col = wid.x;", ""))
