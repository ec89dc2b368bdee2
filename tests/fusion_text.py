"""The WGSL text the reference planner hands a provider, regenerated from a FusionOp list.

`AccelProvider::fused_elementwise(shader, ..)` / `fused_reduction(shader, ..)` receive TEXT produced by
crates/runmat-accelerate/src/fusion.rs (`build_wgsl_shader` :1525-1622, `generate_wgsl_for_output(s)`
:1632-1763, `generate_reduction_wgsl` :1765-2077, expression tables `primitive_expr` :2874-2918 and
`builtin_expr` :2932-3026). The Rust planner is not buildable here, so tests and the bench drive the provider
with the same text emitted by this restatement: same statement grammar (`let tmpN: T = <expr>;`,
`output.data[g] = <expr>;`, `let val: T = <expr>;`), same binding lines, same constants. The provider parses
only that grammar, so what matters is that this file emits exactly the forms the planner emits.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional, Sequence, Union

Operand = Union[int, str]  # int: value id; str: already-formed expression


@dataclass
class FusionOp:
    """FusionOp::{Primitive,Builtin} (fusion.rs:641-652). `inputs` are value ids; `output` a value id."""

    kind: str  # "primitive" | "builtin"
    name: str  # PrimitiveOp name ("Add","Sub","ElemMul","ElemDiv","ElemPow","Neg","UPlus") or builtin name
    inputs: Sequence[int]
    output: int


def _cast_literal(scalar_ty: str, lit: str) -> str:  # fusion.rs:3053-3059
    return f"{scalar_ty}({lit})" if scalar_ty == "f64" else lit


def _expr_is_literal(expr: str, expected: float) -> bool:  # fusion.rs:2920-2930
    t = expr.strip()
    if t.startswith("f64(") and t.endswith(")"):
        t = t[4:-1]
    try:
        return abs(float(t) - expected) <= 2.220446049250313e-16
    except ValueError:
        return False


def primitive_expr(op: str, ins: Sequence[int], exprs: dict) -> Optional[str]:  # fusion.rs:2874-2918
    def binary():
        return exprs[ins[0]], exprs[ins[1]]

    if op == "Add":
        a, b = binary(); return f"({a} + {b})"
    if op == "Sub":
        a, b = binary(); return f"({a} - {b})"
    if op in ("Mul", "ElemMul"):
        a, b = binary(); return f"({a} * {b})"
    if op in ("ElemDiv", "ElemLeftDiv"):
        a, b = binary(); return f"({a} / {b})"
    if op in ("Pow", "ElemPow"):
        a, b = binary()
        if _expr_is_literal(b, 2.0):
            return f"({a} * {a})"
        return f"pow({a}, {b})"
    if op == "Neg":
        return f"(-{exprs[ins[0]]})"
    if op == "UPlus":
        return f"(+{exprs[ins[0]]})"
    return None


_DIRECT = {"sin", "cos", "tan", "asin", "acos", "atan", "sinh", "cosh", "tanh", "exp", "log", "log2", "sqrt", "abs",
           "exp2", "floor", "ceil", "round", "trunc"}


def builtin_expr(name: str, ins: Sequence[int], exprs: dict, scalar_ty: str) -> Optional[str]:  # fusion.rs:2932-3026
    n = name.lower()
    a = exprs[ins[0]]
    if n == "isfinite": return f"isFinite({a})"
    if n == "isinf": return f"isInf({a})"
    if n == "isnan": return f"isNan({a})"
    if n in ("single", "double", "gpuarray"): return a
    if n == "fix": return f"trunc({a})"
    if n == "sign": return f"sign({a})"
    if n == "heaviside":
        zero, half, one = (_cast_literal(scalar_ty, x) for x in ("0.0", "0.5", "1.0"))
        return f"select(select(select({zero}, {one}, ({a} > {zero})), {half}, ({a} == {zero})), {a}, isNan({a}))"
    if n == "mod":
        b = exprs[ins[1]]
        return (f"select(({a} - {b} * floor({a} / {b})), select({b}, {a}, ({a} == 0.0 || sign({a}) == sign({b}))), "
                f"(isInf({b}) && isFinite({a})))")
    if n == "rem":
        b = exprs[ins[1]]
        return f"select(({a} - {b} * trunc({a} / {b})), {a}, (isInf({b}) && isFinite({a})))"
    if n in ("atan2", "hypot", "max", "min"):
        return f"{n}({a}, {exprs[ins[1]]})"
    if n == "pow2":
        return f"exp2({a})" if len(ins) == 1 else None
    if n in ("asinh", "acosh", "atanh"):
        return f"{n}({a})"
    if n in _DIRECT:
        return f"{n}({a})"
    if n == "log10":
        return f"(log({a}) * {_cast_literal(scalar_ty, '0.4342944819032518')})"
    if n == "log1p":
        return f"log({a} + {_cast_literal(scalar_ty, '1.0')})"
    if n == "expm1":
        return f"(exp({a}) - {_cast_literal(scalar_ty, '1.0')})"
    return None


def _prologue(scalar_ty: str, n_inputs: int, output_bindings: str, params_idx: int) -> str:  # fusion.rs:1525-1616
    s = "const MAX_RANK: u32 = 128u;\n"
    s += "struct PackedValue { value: u32, _pad0: u32, _pad1: u32, _pad2: u32 };\n"
    s += "alias PackedArray = array<PackedValue, MAX_RANK>;\n\n"
    s += f"struct Tensor {{ data: array<{scalar_ty}>, }};\n"
    s += "struct Params {\n    len: u32,\n    offset: u32,\n    rank: u32,\n    _pad: u32,\n    out_shape: PackedArray,\n"
    for k in range(n_inputs):
        s += f"    in{k}_shape: PackedArray,\n    in{k}_stride: PackedArray,\n"
    s += "}\n\n"
    if scalar_ty == "f32":
        s += "fn isNan(x: f32) -> bool { let bits = bitcast<u32>(x); return (bits & 0x7f800000u) == 0x7f800000u && (bits & 0x007fffffu) != 0u; }\n"
        s += "fn isFinite(x: f32) -> bool { return (x == x) && (abs(x) < 3.4028234663852886e38); }\n"
        s += "fn isInf(x: f32) -> bool { return (x == x) && !(abs(x) < 3.4028234663852886e38); }\n"
        s += ("fn hypot(a: f32, b: f32) -> f32 {\n    let lo = min(abs(a), abs(b));\n    let hi = max(abs(a), abs(b));\n"
              "    if hi == 0.0 { return 0.0; }\n    if isInf(hi) { return hi; }\n    let r = lo / hi;\n    return hi * sqrt(1.0 + r * r);\n}\n\n")
    else:
        s += "fn isNan(x: f64) -> bool { let bits = bitcast<u64>(x); return (bits & 0x7ff0000000000000u) == 0x7ff0000000000000u && (bits & 0x000fffffffffffffu) != 0u; }\n"
        s += "fn isFinite(x: f64) -> bool { return (x == x) && (abs(x) < f64(1.7976931348623157e308)); }\n"
        s += "fn isInf(x: f64) -> bool { return (x == x) && !(abs(x) < f64(1.7976931348623157e308)); }\n"
        s += ("fn hypot(a: f64, b: f64) -> f64 {\n    let lo = min(abs(a), abs(b));\n    let hi = max(abs(a), abs(b));\n"
              "    if hi == f64(0.0) { return f64(0.0); }\n    if isInf(hi) { return hi; }\n    let r = lo / hi;\n    return hi * sqrt(f64(1.0) + r * r);\n}\n\n")
    for k in range(n_inputs):
        s += f"@group(0) @binding({k}) var<storage, read> input{k}: Tensor;\n"
    s += output_bindings
    s += f"@group(0) @binding({params_idx}) var<uniform> params: Params;\n\n"
    s += "@compute @workgroup_size(@WG@)\nfn main(@builtin(global_invocation_id) gid: vec3<u32>) {\n"
    s += "    let idx = gid.x;\n    if (idx >= params.len) { return; }\n    let g = idx + params.offset;\n"
    s += ("    var coord: array<u32, MAX_RANK>;\n    var tmp: u32 = g;\n    var d: u32 = 0u;\n    loop { if d >= params.rank { break; } "
          "let dim = params.out_shape[d].value; if dim == 0u { coord[d] = 0u; } else { coord[d] = tmp % dim; tmp = tmp / dim; } d = d + 1u; }\n")
    for k in range(n_inputs):
        s += (f"    var i{k}: u32 = 0u; d = 0u; loop {{ if d >= params.rank {{ break; }} let sd = params.in{k}_shape[d].value; "
              f"let st = params.in{k}_stride[d].value; let c = select(coord[d], 0u, sd == 1u); i{k} = i{k} + c * st; d = d + 1u; }}\n")
    return s


def elementwise_wgsl(input_ids: Sequence[int], ops: Sequence[FusionOp], output_ids: Sequence[int], scalar_ty: str = "f64") -> str:
    """generate_wgsl_for_output / generate_wgsl_for_outputs (fusion.rs:1632-1763)."""
    exprs = {vid: f"input{k}.data[i{k}]" for k, vid in enumerate(input_ids)}
    body = ""
    for node_idx, op in enumerate(ops):
        tmp = f"tmp{node_idx}"
        e = primitive_expr(op.name, op.inputs, exprs) if op.kind == "primitive" else builtin_expr(op.name, op.inputs, exprs, scalar_ty)
        if e is None:
            raise ValueError(f"unsupported fusion op {op}")
        body += f"    let {tmp}: {scalar_ty} = {e};\n"
        exprs[op.output] = tmp
    n_in = len(input_ids)
    if len(output_ids) == 1:
        bindings = f"@group(0) @binding({n_in}) var<storage, read_write> output: Tensor;\n"
        writes = f"    output.data[g] = {exprs[output_ids[0]]};\n"
        params_idx = n_in + 1
    else:
        bindings = "".join(f"@group(0) @binding({n_in + k}) var<storage, read_write> output{k}: Tensor;\n" for k in range(len(output_ids)))
        writes = "".join(f"    output{k}.data[g] = {exprs[o]};\n" for k, o in enumerate(output_ids))
        params_idx = n_in + len(output_ids)
    return _prologue(scalar_ty, n_in, bindings, params_idx) + body + writes + "}\n"


def reduction_wgsl(input_ids: Sequence[int], ops: Sequence[FusionOp], data_id: int, axis: int = 0, omitnan: bool = False,
                   mean: bool = False, scalar_ty: str = "f64", const_values: Optional[dict] = None) -> str:
    """generate_reduction_wgsl (fusion.rs:1765-2077): `v`, `v1`.. operands, OMITNAN const, axis-specific loop."""
    exprs = {input_ids[0]: "v"}
    for k, vid in enumerate(input_ids[1:], start=1):
        exprs[vid] = f"v{k}"
    for vid, val in (const_values or {}).items():
        exprs[vid] = f"f64({_fmt_num(val)})" if scalar_ty == "f64" else repr(float(val))
    for op in ops:
        e = primitive_expr(op.name, op.inputs, exprs) if op.kind == "primitive" else builtin_expr(op.name, op.inputs, exprs, scalar_ty)
        if e is None:
            raise ValueError(f"unsupported fusion op {op}")
        exprs[op.output] = e
    val = exprs[data_id]
    n_in = len(input_ids)
    s = f"struct Tensor {{ data: array<{scalar_ty}>, }};\n"
    s += "struct MParams { nrows: u32, ncols: u32, ld: u32, flags: u32 }\n\n"
    for k in range(n_in):
        s += f"@group(0) @binding({k}) var<storage, read> input{k}: Tensor;\n"
    s += f"@group(0) @binding({n_in}) var<storage, read_write> output: Tensor;\n"
    s += f"@group(0) @binding({n_in + 1}) var<uniform> params: MParams;\n\n"
    s += f"var<workgroup> tile: array<{scalar_ty}, @WG@u>;\n\n"
    s += f"const OMITNAN: bool = {'true' if omitnan else 'false'};\n\n"
    dim = "params.nrows" if axis == 0 else "params.ncols"
    if mean:
        post = f"(1.0 / f64(f32({dim})))" if scalar_ty == "f64" else f"(1.0 / f32({dim}))"
    else:
        post = "f64(1.0)" if scalar_ty == "f64" else "1.0"
    s += f"fn isNanF(x: {scalar_ty}) -> bool {{ return x != x; }}\n"
    if scalar_ty == "f64":
        s += "fn canonicalNan() -> f64 {\n  var bits: u64 = 0x7ff8000000000000u;\n  return bitcast<f64>(bits);\n}\n\n"
    else:
        s += "fn canonicalNan() -> f32 {\n  var bits: u32 = 0x7fc00000u;\n  return bitcast<f32>(bits);\n}\n\n"
    s += "@compute @workgroup_size(@WG@)\n"
    s += "fn main(@builtin(local_invocation_id) lid: vec3<u32>, @builtin(workgroup_id) wid: vec3<u32>) {\n"
    acc0 = "f64(0.0" if scalar_ty == "f64" else "0.0"
    if axis == 0:
        s += "  let col = wid.x;\n  if (col >= params.ncols) { return; }\n"
        s += f"  var acc: {scalar_ty} = {acc0};\n"
        s += "  var saw_nan: bool = false;\n  var r = lid.x;\n  while (r < params.nrows) {\n"
        s += "    let v = input0.data[ (col * params.nrows) + r ];\n"
        for k in range(1, n_in):
            s += f"    let v{k} = input{k}.data[ (col * params.nrows) + r ];\n"
        s += (f"    let val: {scalar_ty} = {val};\n    if (OMITNAN) {{ if (!isNanF(val)) {{ acc = acc + val; }} }} else {{ if (isNanF(val)) "
              "{ saw_nan = true; } else { acc = acc + val; } }\n")
        s += "    r += @WG@u;\n  }\n"
        out_idx = "col"
    else:
        s += "  let row = wid.x;\n  if (row >= params.ncols) { return; }\n"
        s += f"  var acc: {scalar_ty} = {acc0};\n"
        s += "  var saw_nan: bool = false;\n  var c = lid.x;\n  while (c < params.nrows) {\n"
        s += "    let v = input0.data[ row + (c * params.ncols) ];\n"
        for k in range(1, n_in):
            s += f"    let v{k} = input{k}.data[ row + (c * params.ncols) ];\n"
        s += (f"    let val: {scalar_ty} = {val};\n    if (OMITNAN) {{ if (!isNanF(val)) {{ acc = acc + val; }} }} else {{ if (isNanF(val)) "
              "{ saw_nan = true; } else { acc = acc + val; } }\n")
        s += "    c += @WG@u;\n  }\n"
        out_idx = "row"
    s += "  if (!OMITNAN && saw_nan) { acc = canonicalNan(); }\n  tile[lid.x] = acc;\n  workgroupBarrier();\n"
    s += ("  var off = (@WG@u) / 2u;\n  loop { if (off == 0u) { break; } if (lid.x < off) {\n    let a = tile[lid.x]; let b = tile[lid.x + off];\n"
          "    tile[lid.x] = a + b;\n  } workgroupBarrier(); off = off / 2u; }\n")
    s += f"  if (lid.x == 0u) {{ output.data[{out_idx}] = tile[0u] * {post}; }}\n}}\n"
    return s


def _fmt_num(v: float) -> str:
    """Rust `{}` formatting of an f64 (integers print without a fraction: 2.0 -> "2")."""
    f = float(v)
    if f == int(f) and abs(f) < 1e16:
        return str(int(f))
    return repr(f)


# ---- the headline programs ------------------------------------------------------------------------------------------
def sin_mul_add_program():
    """C = sin(A) .* B + 1: ops [Builtin sin(A)->t0, ElemMul(t0,B)->t1, Add(t1,1)->C]; the scalar `1` is a
    1-element input tensor (fusion_exec.rs:305-326). Returns (input_ids, ops, output_id)."""
    A, B, ONE, T0, T1, Cc = 0, 1, 2, 10, 11, 12
    ops = [FusionOp("builtin", "sin", [A], T0), FusionOp("primitive", "ElemMul", [T0, B], T1), FusionOp("primitive", "Add", [T1, ONE], Cc)]
    return [A, B, ONE], ops, Cc


def sin_mul_add_wgsl(scalar_ty: str = "f64") -> str:
    ins, ops, out = sin_mul_add_program()
    return elementwise_wgsl(ins, ops, [out], scalar_ty)


def sum_sin_mul_add_wgsl(scalar_ty: str = "f64") -> str:
    """sum(sin(A).*B + 1, 'all') as ONE reduction program (16 B/elem instead of the reference's 32 B/elem
    elementwise+reduce pair). The constant folds as a literal, as `const_values` do."""
    A, B, ONE, T0, T1, Cc = 0, 1, 2, 10, 11, 12
    ops = [FusionOp("builtin", "sin", [A], T0), FusionOp("primitive", "ElemMul", [T0, B], T1), FusionOp("primitive", "Add", [T1, ONE], Cc)]
    return reduction_wgsl([A, B], ops, Cc, axis=0, scalar_ty=scalar_ty, const_values={ONE: 1.0})
