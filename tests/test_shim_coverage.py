"""Boundary completeness: the C ABI header, the Rust shim and the reference trait must agree.

  * every function include/rm_accel.h declares is either an explicit EXTENSION (no counterpart in the single-device trait:
    lifecycle, measurement, multi-GPU exchange, f32/async transfer helpers, debug hooks) or is declared in the shim's
    `extern "C"` block AND called from `impl AccelProvider for CudaProvider`;
  * every trait method listed in TRAIT_TO_C is implemented by the shim, and (when /root/reference is present, i.e. in the build
    container; the GPU box does not carry it) exists in `runmat_accelerate_api::AccelProvider` with the same parameter count.
No compute, no GPU."""
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = (ROOT / "include" / "rm_accel.h").read_text()
SHIM = (ROOT / "shim" / "cuda_provider.rs").read_text()
TRAIT_SRC = Path("/root/reference/crates/runmat-accelerate-api/src/lib.rs")

# C entry points that have no method on the reference trait (kept out of the forwarding requirement on purpose)
EXTENSIONS = {
    # lifecycle / registry glue used by CudaProvider::new / Drop and by hosts that embed the library directly
    "rm_abi_version", "rm_device_info_struct", "rm_device_id", "rm_provider_precision", "rm_synchronize", "rm_host_sync_count",
    "rm_get_stream", "rm_set_stream", "rm_device_ptr", "rm_copy_to_device", "rm_device_pci_bus_id", "rm_live_buffers", "rm_live_bytes",
    # transfer helpers for f32 / pipelined hosts
    "rm_upload_f32", "rm_download_f32", "rm_download_async", "rm_pinned_alloc", "rm_pinned_free",
    # multi-GPU exchange (the trait is single-device)
    "rm_comm_unique_id", "rm_comm_init", "rm_comm_world_size", "rm_comm_allreduce_sum", "rm_comm_fence", "rm_comm_p2p_export",
    "rm_comm_p2p_connect", "rm_comm_p2p_connected", "rm_comm_p2p_error", "rm_fused_reduction_allreduce", "rm_stochastic_evolution_sharded",
    "rm_payoff_partial_sum",
    # measurement / tuning / debug
    "rm_timer_begin", "rm_timer_end_ms", "rm_flush_l2", "rm_set_matmul_engine", "rm_debug_ozaki_stats", "rm_debug_device_flags", "rm_set_launch_overlap", "rm_get_rng_state",
    # the named per-op images are reached through the generic dispatchers (rm_elem_binary / rm_unary / rm_scalar_op_apply)
    "rm_elem_add", "rm_elem_mul", "rm_elem_max", "rm_elem_min", "rm_elem_sub", "rm_elem_div", "rm_elem_pow", "rm_elem_hypot", "rm_elem_atan2",
    "rm_unary_sin", "rm_unary_cos", "rm_unary_tan", "rm_unary_tanh", "rm_unary_exp", "rm_unary_log", "rm_unary_sqrt", "rm_unary_abs",
    "rm_unary_floor", "rm_unary_round", "rm_scalar_add", "rm_scalar_sub", "rm_scalar_mul", "rm_scalar_div", "rm_scalar_rsub", "rm_scalar_rdiv",
    "rm_scalar_max", "rm_scalar_min",
}

# trait method -> C entry point it forwards to (AccelProvider, crates/runmat-accelerate-api/src/lib.rs:1386-3152)
TRAIT_TO_C = {
    "upload": "rm_upload", "download": "rm_download", "free": "rm_free", "device_info": "rm_device_info_string", "read_scalar": "rm_read_scalar",
    "zeros": "rm_zeros", "ones": "rm_ones", "ones_like": "rm_ones", "fill": "rm_fill", "eye": "rm_eye", "linspace": "rm_linspace", "reshape": "rm_reshape",
    "transpose": "rm_transpose", "permute": "rm_permute", "repmat": "rm_repmat", "cat": "rm_cat", "gather_linear": "rm_gather_linear",
    "scatter_linear": "rm_scatter_linear", "find": "rm_find", "scatter_column": "rm_scatter_column", "scatter_row": "rm_scatter_row",
    "sub2ind": "rm_sub2ind", "ind2sub": "rm_ind2sub",
    "random_uniform": "rm_random_uniform", "random_uniform_like": "rm_random_uniform", "random_normal": "rm_random_normal",
    "random_normal_like": "rm_random_normal", "set_rng_state": "rm_set_rng_state", "stochastic_evolution": "rm_stochastic_evolution",
    "fused_elementwise": "rm_fused_elementwise", "fused_elementwise_multi": "rm_fused_elementwise_multi", "fused_reduction": "rm_fused_reduction",
    "fused_cache_counters": "rm_fused_cache_counters",
    "reduce_sum": "rm_reduce_sum", "reduce_sum_dim": "rm_reduce_sum_dim", "reduce_prod": "rm_reduce_prod", "reduce_mean": "rm_reduce_mean",
    "reduce_mean_dim": "rm_reduce_mean_dim", "reduce_mean_nd": "rm_reduce_mean_nd", "reduce_moments_nd": "rm_reduce_moments_nd",
    "reduce_max": "rm_reduce_max", "reduce_min": "rm_reduce_min", "reduce_max_dim": "rm_reduce_max_dim", "reduce_min_dim": "rm_reduce_min_dim",
    "default_reduction_workgroup_size": "rm_default_reduction_workgroup_size", "two_pass_threshold": "rm_two_pass_threshold",
    "matmul": "rm_matmul", "matmul_epilogue": "rm_matmul_epilogue_apply", "syrk": "rm_syrk", "matmul_power_step": "rm_matmul_power_step",
    "covariance": "rm_covariance", "diag_extract": "rm_diag_extract", "mldivide": "rm_mldivide", "mrdivide": "rm_mrdivide", "linsolve": "rm_linsolve",
    "image_normalize": "rm_image_normalize", "imfilter": "rm_imfilter", "conv2d": "rm_conv2d",
    "telemetry_snapshot": "rm_telemetry_snapshot", "reset_telemetry": "rm_reset_telemetry", "warmup": "rm_warmup",
}
# trait methods served by the generic dispatchers, with the op code the shim must pass (include/rm_accel.h enums)
BINARY = {"elem_add": 0, "elem_sub": 1, "elem_mul": 2, "elem_div": 3, "elem_pow": 4, "elem_max": 5, "elem_min": 6, "elem_hypot": 7, "elem_atan2": 8,
          "elem_ge": 11, "elem_le": 12, "elem_lt": 13, "elem_gt": 14, "elem_eq": 15, "elem_ne": 16}
UNARY = {"unary_sin": 0, "unary_cos": 1, "unary_tan": 2, "unary_asin": 3, "unary_acos": 4, "unary_atan": 5, "unary_sinh": 6, "unary_cosh": 7, "unary_tanh": 8,
         "unary_asinh": 9, "unary_acosh": 10, "unary_atanh": 11, "unary_exp": 12, "unary_expm1": 13, "unary_log": 14, "unary_log2": 15, "unary_log10": 16,
         "unary_log1p": 17, "unary_sqrt": 18, "unary_abs": 19, "unary_sign": 20, "unary_floor": 21, "unary_ceil": 22, "unary_round": 23, "unary_fix": 24,
         "unary_pow2": 26, "unary_heaviside": 27, "unary_single": 28, "unary_double": 29, "logical_isnan": 30, "logical_isinf": 31, "logical_isfinite": 32,
         "map_nan_to_zero": 33, "not_nan_mask": 34, "unary_erf": 35, "unary_gamma": 36, "unary_gammaln": 37}
SCALAR = {"scalar_add": 0, "scalar_sub": 1, "scalar_mul": 2, "scalar_div": 3, "scalar_rsub": 4, "scalar_rdiv": 5, "scalar_max": 6, "scalar_min": 7}


def header_functions():
    return set(re.findall(r"\b(rm_[a-z0-9_]+)\s*\(", HEADER))


def shim_extern_block():
    m = re.search(r'extern "C" \{(.*?)\n\}', SHIM, re.S)
    return set(re.findall(r"fn (rm_[a-z0-9_]+)\s*\(", m.group(1)))


def shim_impl():
    body = SHIM[SHIM.index("\nimpl AccelProvider for CudaProvider"):]
    return body, dict((m.group(1), m.start()) for m in re.finditer(r"\n    fn (\w+)", body))


def fn_text(body, offsets, name):
    start = offsets[name]
    later = [o for o in offsets.values() if o > start]
    return body[start:min(later) if later else len(body)]


def param_count(sig: str) -> int:
    """Parameters after `self` in `fn name<..>( &self, a: T, b: U ) -> R` (generic-aware comma split)."""
    inner = sig[sig.index("(") + 1:]
    depth, cur, parts = 0, "", []
    for ch in inner:
        if ch in "<([":
            depth += 1
        elif ch in ">)]":
            if ch == ")" and depth == 0:
                break
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur)
            cur = ""
        else:
            cur += ch
    parts.append(cur)
    parts = [p.strip() for p in parts if p.strip()]
    return len([p for p in parts if not re.fullmatch(r"&?('\w+\s+)?(mut\s+)?self", p)])


def enum_values(name):
    m = re.search(r"typedef enum %s \{(.*?)\}" % name, HEADER, re.S)
    body = re.sub(r"/\*.*?\*/", "", m.group(1), flags=re.S)
    out, v = {}, 0
    for tok in body.replace("\n", " ").split(","):
        tok = tok.strip()
        if not tok:
            continue
        if "=" in tok:
            tok, val = [x.strip() for x in tok.split("=")]
            v = int(val)
        out[tok] = v
        v += 1
    return out


def test_every_exported_entry_point_is_forwarded_or_an_extension():
    hdr, ext = header_functions(), shim_extern_block()
    body, offsets = shim_impl()
    assert not (EXTENSIONS - hdr), f"stale EXTENSIONS entries: {sorted(EXTENSIONS - hdr)}"
    assert not (ext - hdr), f"shim binds symbols the header does not declare: {sorted(ext - hdr)}"
    need = hdr - EXTENSIONS
    assert not (need - ext), f"exported but not bound by the shim: {sorted(need - ext)}"
    uses = SHIM[SHIM.index("pub struct CudaProvider"):]   # everything after the extern block
    unused = [f for f in sorted(need) if not re.search(r"\b%s\b" % f, uses)]
    assert not unused, f"bound but never called from the provider impl: {unused}"
    assert set(TRAIT_TO_C.values()) <= hdr


def test_trait_methods_forward_to_the_mapped_entry_points_with_the_right_op_codes():
    body, offsets = shim_impl()
    for method, cfn in TRAIT_TO_C.items():
        assert method in offsets, f"shim does not implement AccelProvider::{method}"
        text = fn_text(body, offsets, method)
        helpers = {"rm_zeros": "self.zeros(", "rm_fill": "self.fill("}
        assert re.search(r"\b%s\b" % cfn, text) or helpers.get(cfn, "\0") in text, f"{method} does not call {cfn}"
    bins, uns, scs = enum_values("rm_binary_op"), enum_values("rm_unary_op"), enum_values("rm_scalar_op")
    for table, helper, enum, prefix in ((BINARY, "binary", bins, "RM_BIN_"), (UNARY, "unary", uns, "RM_UN_"), (SCALAR, "scalar", scs, "RM_SC_")):
        for method, code in table.items():
            assert method in offsets, f"shim does not implement AccelProvider::{method}"
            m = re.search(r"self\.%s\((\d+)," % helper, fn_text(body, offsets, method))
            assert m and int(m.group(1)) == code, f"{method}: expected op code {code}"
            cname = method.split("_", 1)[1].upper()
            cname = {"ISNAN": "ISNAN", "ISINF": "ISINF", "ISFINITE": "ISFINITE", "NAN_TO_ZERO": "NAN_TO_ZERO", "NAN_MASK": "NOT_NAN_MASK"}.get(cname, cname)
            if method == "map_nan_to_zero":
                cname = "NAN_TO_ZERO"
            assert enum[prefix + cname] == code, f"{method}: header says {prefix}{cname} = {enum[prefix + cname]}, shim passes {code}"


@pytest.mark.skipif(not TRAIT_SRC.exists(), reason="reference sources are only present in the build container")
def test_shim_methods_exist_on_the_reference_trait_with_the_same_arity():
    src = TRAIT_SRC.read_text()
    start = src.index("pub trait AccelProvider")
    i = src.index("{", start)
    depth = 0
    for j in range(i, len(src)):
        depth += src[j] == "{"
        depth -= src[j] == "}"
        if depth == 0:
            break
    trait = src[i:j]
    trait_sigs = {m.group(1): m.group(0) for m in re.finditer(r"\n    fn (\w+)[^{;]*", trait)}
    body, offsets = shim_impl()
    for name in offsets:
        assert name in trait_sigs, f"shim implements {name}, which AccelProvider does not declare"
        sig = re.match(r"\n    fn \w+[^{]*", fn_text(body, offsets, name)).group(0)
        assert param_count(sig) == param_count(trait_sigs[name]), f"{name}: shim takes {param_count(sig)} parameters, the trait {param_count(trait_sigs[name])}"
    # every hot-path trait method the library can serve is implemented
    for name in list(TRAIT_TO_C) + list(BINARY) + list(UNARY) + list(SCALAR):
        assert name in trait_sigs, f"{name} is not a method of the reference trait"
