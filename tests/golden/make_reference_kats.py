"""Writes tests/golden/reference_kats.json: literal known-answer vectors transcribed from the reference's own
unit tests (paths relative to /root/reference/crates). The reference is Rust and cannot be executed here, so
these are the literals its tests assert, not outputs of running it. Re-run to regenerate the JSON."""
import json
from pathlib import Path

NAN, INF = "nan", "inf"
kats = {
    "elem_binary": [
        {"src": "runmat-runtime/src/builtins/math/elementwise/times.rs:932-943 times_matrix_scalar", "op": "mul",
         "a": {"shape": [2, 2], "data": [1, 2, 3, 4]}, "b": {"shape": [1, 1], "data": [2]},
         "out": {"shape": [2, 2], "data": [2, 4, 6, 8]}},
        {"src": "runmat-runtime/src/builtins/math/elementwise/times.rs:947-962 times_row_column_broadcast", "op": "mul",
         "a": {"shape": [3, 1], "data": [1, 2, 3]}, "b": {"shape": [1, 3], "data": [10, 20, 30]},
         "out": {"shape": [3, 3], "data": [10, 20, 30, 20, 40, 60, 30, 60, 90]}},
    ],
    # MATLAB-semantics values of mod/rem for the inputs of mod_real_parity_matrix_matches_expected_values /
    # rem_real_parity_matrix_matches_expected_values (runmat-vm/tests/fusion_gpu.rs:3056-3222), evaluated by
    # hand from mod_real_expected/rem_real_expected (:2965-3030); b==0 -> NaN is the reference's choice.
    "mod": [[-5.5, 2, 0.5], [-1.25, 2, 0.75], [5.5, 2, 1.5], [6, 2, 0], [6, 3, 0], [6, 4, 2], [-1.25, 3, 1.75], [5.5, 4, 1.5],
            [6, 0, NAN], [6, INF, 6], [NAN, 2, NAN], [5, INF, 5], [INF, 2, NAN], [4, 0, NAN],
            [5, "-inf", "-inf"], [-5, "-inf", -5], [-5, INF, INF], [0, "-inf", 0]],
    "rem": [[-5.5, 2, -1.5], [-1.25, 2, -1.25], [5.5, 2, 1.5], [6, 2, 0], [6, 3, 0], [6, 4, 2], [-1.25, 3, -1.25], [5.5, 4, 1.5],
            [6, 0, NAN], [6, INF, 6], [NAN, 2, NAN], [5, INF, 5], [INF, 2, NAN], [4, 0, NAN]],
    "sum": [
        {"src": "runmat-runtime/src/builtins/math/reduction/sum.rs:1445-1456 sum_matrix_default_dimension",
         "a": {"shape": [2, 3], "data": [1, 4, 2, 5, 3, 6]}, "dims": [0], "omitnan": False, "out": {"shape": [1, 3], "data": [5, 7, 9]}},
        {"src": "sum.rs:1460-1472 sum_matrix_dimension_two",
         "a": {"shape": [2, 3], "data": [1, 4, 2, 5, 3, 6]}, "dims": [1], "omitnan": False, "out": {"shape": [2, 1], "data": [6, 15]}},
        {"src": "sum.rs:1476-1481 sum_all_dimension",
         "a": {"shape": [2, 3], "data": [1, 2, 3, 4, 5, 6]}, "dims": [0, 1], "omitnan": False, "out": {"shape": [1, 1], "data": [21]}},
        {"src": "sum.rs:1485-1500 sum_vecdim_multiple_axes",
         "a": {"shape": [3, 4, 2], "data": list(range(1, 25))}, "dims": [0, 2], "omitnan": False,
         "out": {"shape": [1, 4, 1], "data": [48, 66, 84, 102]}},
        {"src": "sum.rs:1504-1508 sum_with_omit_nan_default_dimension",
         "a": {"shape": [3, 1], "data": [1, NAN, 3]}, "dims": [0], "omitnan": True, "out": {"shape": [1, 1], "data": [4]}},
        {"src": "sum.rs:1512-1519 sum_with_include_nan_propagates",
         "a": {"shape": [3, 1], "data": [1, NAN, 3]}, "dims": [0], "omitnan": False, "out": {"shape": [1, 1], "data": [NAN]}},
    ],
    "mean": [
        {"src": "runmat-runtime/src/builtins/math/reduction/mean.rs:1695-1705 mean_matrix_default_dimension",
         "a": {"shape": [2, 3], "data": [1, 4, 2, 5, 3, 6]}, "dims": [0], "omitnan": False, "out": {"shape": [1, 3], "data": [2.5, 3.5, 4.5]}},
        {"src": "mean.rs:1709-1720 mean_matrix_dimension_two",
         "a": {"shape": [2, 3], "data": [1, 4, 2, 5, 3, 6]}, "dims": [1], "omitnan": False, "out": {"shape": [2, 1], "data": [2, 5]}},
        {"src": "mean.rs:1724-1732 mean_with_omit_nan_default_dimension (tol 1e-12)",
         "a": {"shape": [3, 1], "data": [1, NAN, 5]}, "dims": [0], "omitnan": True, "out": {"shape": [1, 1], "data": [3]}},
        {"src": "mean.rs:1736-1744 mean_with_omit_nan_all_nan_returns_nan",
         "a": {"shape": [2, 1], "data": [NAN, NAN]}, "dims": [0], "omitnan": True, "out": {"shape": [1, 1], "data": [NAN]}},
        {"src": "mean.rs:1748-1756 mean_with_include_nan_propagates_nan",
         "a": {"shape": [3, 1], "data": [1, NAN, 3]}, "dims": [0], "omitnan": False, "out": {"shape": [1, 1], "data": [NAN]}},
    ],
    "prod": [
        {"src": "runmat-runtime/src/builtins/math/reduction/prod.rs:1248-1258 prod_matrix_default_dimension",
         "a": {"shape": [2, 3], "data": [1, 4, 2, 5, 3, 6]}, "dims": [0], "out": {"shape": [1, 3], "data": [4, 10, 18]}},
        {"src": "prod.rs:1262-1273 prod_matrix_dimension_two",
         "a": {"shape": [2, 3], "data": [1, 4, 2, 5, 3, 6]}, "dims": [1], "out": {"shape": [2, 1], "data": [6, 120]}},
        {"src": "prod.rs:1277-1282 prod_all_dimension",
         "a": {"shape": [2, 3], "data": [1, 2, 3, 4, 5, 6]}, "dims": [0, 1], "out": {"shape": [1, 1], "data": [720]}},
        {"src": "prod.rs:1286-1301 prod_vecdim_multiple_axes",
         "a": {"shape": [3, 4, 2], "data": list(range(1, 25))}, "dims": [0, 2], "out": {"shape": [1, 4, 1], "data": [16380, 587520, 4021920, 16030080]}},
    ],
    "max_dim": [
        {"src": "runmat-runtime/src/builtins/math/reduction/max.rs:2471-2477 max_vector_with_indices",
         "a": {"shape": [3, 1], "data": [3, 1, 5]}, "dim": 0, "values": [5], "indices": [3]},
        {"src": "max.rs:2608-2625 max_matrix_default_dimension",
         "a": {"shape": [2, 3], "data": [3, 4, 1, 2, 5, 6]}, "dim": 0, "values": [4, 2, 6], "indices": [2, 2, 2]},
    ],
    "find": [
        {"src": "runmat-runtime/src/builtins/array/indexing/find.rs:885-895 find_linear_indices_basic",
         "a": {"shape": [2, 3], "data": [0, 4, 0, 7, 0, 9]}, "limit": None, "direction": "first", "linear": [2, 4, 6]},
        {"src": "find.rs:912-922 find_limited_first", "a": {"shape": [1, 5], "data": [0, 3, 5, 0, 8]}, "limit": 2, "direction": "first", "linear": [2, 3]},
        {"src": "find.rs:926-936 find_last_single", "a": {"shape": [1, 6], "data": [1, 0, 0, 6, 0, 2]}, "limit": 1, "direction": "last", "linear": [6]},
        {"src": "find.rs:990-1006 find_multi_output_rows_cols_values", "a": {"shape": [2, 3], "data": [0, 2, 3, 0, 0, 6]}, "limit": None,
         "direction": "first", "linear": [2, 3, 6], "rows": [2, 1, 2], "cols": [1, 2, 3], "vals": [2, 3, 6]},
    ],
    "sub2ind": [
        {"src": "runmat-runtime/src/builtins/array/indexing/sub2ind.rs:460-465 converts_scalar_indices", "dims": [3, 4],
         "subs": [{"shape": [1, 1], "data": [2]}, {"shape": [1, 1], "data": [3]}], "out": {"shape": [1, 1], "data": [8]}},
        {"src": "sub2ind.rs:496-511 broadcasts_scalars_over_vectors", "dims": [3, 4],
         "subs": [{"shape": [3, 1], "data": [1, 2, 3]}, {"shape": [1, 1], "data": [4]}], "out": {"shape": [3, 1], "data": [10, 11, 12]}},
        {"src": "sub2ind.rs:515-532 handles_three_dimensions", "dims": [2, 3, 4],
         "subs": [{"shape": [1, 2], "data": [1, 1]}, {"shape": [1, 2], "data": [2, 3]}, {"shape": [1, 2], "data": [1, 2]}],
         "out": {"shape": [1, 2], "data": [3, 11]}},
        {"src": "sub2ind.rs:536-549 rejects_out_of_range_subscripts", "dims": [3, 4],
         "subs": [{"shape": [1, 1], "data": [4]}, {"shape": [1, 1], "data": [1]}], "error": "Index exceeds"},
    ],
    "ind2sub": [
        {"src": "runmat-runtime/src/builtins/array/indexing/ind2sub.rs:376-388 recovers_tensor_indices", "dims": [3, 4],
         "idx": {"shape": [1, 1], "data": [8]}, "out": [[2], [3]]},
        {"src": "ind2sub.rs:406-432 handles_vector_indices", "dims": [3, 5], "idx": {"shape": [1, 3], "data": [7, 8, 9]},
         "out": [[1, 2, 3], [3, 3, 3]]},
        {"src": "ind2sub.rs:460-484 recovers_three_dimensional_indices", "dims": [2, 3, 4], "idx": {"shape": [1, 2], "data": [3, 11]},
         "out": [[1, 1], [2, 3], [1, 2]]},
    ],
    "permute": [
        {"src": "runmat-runtime/src/builtins/array/shape/permute.rs:619-631 permute_swaps_dims (shape only in the reference test)",
         "a": {"shape": [2, 3, 4], "data": list(range(1, 25))}, "order": [2, 1, 3], "out_shape": [3, 2, 4]},
        {"src": "permute.rs:635-645 permute_adds_trailing_dimension", "a": {"shape": [1, 3], "data": [1, 2, 3]}, "order": [2, 1, 3], "out_shape": [3, 1, 1]},
    ],
    "repmat": [
        {"src": "runmat-runtime/src/builtins/array/shape/repmat.rs:755-781 repeats_matrix_with_vector_reps",
         "a": {"shape": [2, 2], "data": [1, 3, 2, 4]}, "reps": [2, 3],
         "out": {"shape": [4, 6], "data": [1, 3, 1, 3, 2, 4, 2, 4] * 3}},
        {"src": "repmat.rs:812-844 repmat_high_dim_numeric (out[i,j,k] = base[j%3 + 3*(k%2)])",
         "a": {"shape": [1, 3, 2], "data": [0, 1, 2, 3, 4, 5]}, "reps": [2, 1, 3],
         "out": {"shape": [2, 3, 6], "data": [float((j % 3) + 3 * (k % 2)) for k in range(6) for j in range(3) for _ in range(2)]}},
    ],
    "cat": [
        {"src": "runmat-runtime/src/builtins/array/shape/cat.rs:1305-1320 cat_numeric_rows", "dim": 1,
         "inputs": [{"shape": [2, 2], "data": [1, 3, 2, 4]}, {"shape": [2, 2], "data": [5, 7, 6, 8]}],
         "out": {"shape": [4, 2], "data": [1, 3, 5, 7, 2, 4, 6, 8]}},
    ],
    "eye": [
        {"src": "runmat-runtime/src/builtins/array/creation/eye.rs:597-617 eye_rectangular_from_two_dims", "rows": 2, "cols": 4,
         "out": {"shape": [2, 4], "data": [1, 0, 0, 1, 0, 0, 0, 0]}},
    ],
    "matmul": [
        {"src": "runmat-runtime/src/builtins/math/linalg/ops/mtimes.rs:495-507 matrix_product_matches_expected",
         "a": {"shape": [2, 3], "data": [1, 4, 2, 5, 3, 6]}, "b": {"shape": [3, 2], "data": [7, 9, 11, 8, 10, 12]},
         "out": {"shape": [2, 2], "data": [58, 139, 64, 154]}},
    ],
    "matmul_epilogue": [
        {"src": "runmat-accelerate/tests/matmul_epilogue.rs:24-110 matmul_epilogue_row_col_alpha_beta (tol 1e-9; expected = (alpha*(A*B)+beta).*row.*col)",
         "a": {"shape": [3, 2], "data": [1, 2, 3, 4, 5, 6]}, "b": {"shape": [2, 4], "data": [1, 3, 5, 7, 2, 4, 6, 8]},
         "alpha": 1.25, "beta": -0.5, "row_scale": [2.0, 0.5, 1.0], "col_scale": [1.0, 2.0, 0.25, 1.5]},
    ],
    "imfilter": [
        {"src": "runmat-runtime/src/builtins/image/filters/imfilter.rs:803-811 same_padding_default_zero",
         "img": {"shape": [2, 2], "data": [1, 3, 2, 4]}, "ker": {"shape": [3, 3], "data": [1] * 9}, "opts": {}, "out": {"shape": [2, 2], "data": [10, 10, 10, 10]}},
        {"src": "imfilter.rs:814-829 replicate_padding",
         "img": {"shape": [2, 2], "data": [1, 3, 2, 4]}, "ker": {"shape": [3, 3], "data": [1] * 9}, "opts": {"padding": "replicate"},
         "out": {"shape": [2, 2], "data": [18, 24, 21, 27]}},
        {"src": "imfilter.rs:832-847 full_output_matches_expected_size",
         "img": {"shape": [2, 2], "data": [1, 3, 2, 4]}, "ker": {"shape": [2, 2], "data": [1, 3, 2, 4]}, "opts": {"shape": "full"},
         "out": {"shape": [3, 3], "data": [4, 14, 6, 11, 30, 11, 6, 14, 4]}},
        {"src": "imfilter.rs:850-862 valid_output_respects_kernel_size",
         "img": {"shape": [2, 2], "data": [1, 2, 3, 4]}, "ker": {"shape": [2, 2], "data": [1] * 4}, "opts": {"shape": "valid"},
         "out": {"shape": [1, 1], "data": [10]}},
        {"src": "imfilter.rs:897-912 circular_padding_wraps_indices",
         "img": {"shape": [2, 2], "data": [1, 2, 3, 4]}, "ker": {"shape": [2, 2], "data": [0, 1, 1, 0]}, "opts": {"padding": "circular"},
         "out": {"shape": [2, 2], "data": [5, 5, 5, 5]}},
        {"src": "imfilter.rs:915-947 gpu_fallback_uses_provider_upload",
         "img": {"shape": [2, 2], "data": [1, 4, 2, 5]}, "ker": {"shape": [2, 2], "data": [1, 1, 1, 1]}, "opts": {},
         "out": {"shape": [2, 2], "data": [1, 5, 3, 12]}},
        {"src": "imfilter.rs:967-984 doc_example_convolution_same_matches_expected",
         "img": {"shape": [3, 2], "data": [1, 2, 3, 4, 5, 6]}, "ker": {"shape": [2, 2], "data": [1, 3, 2, 4]}, "opts": {"mode": "conv"},
         "out": {"shape": [3, 2], "data": [1, 5, 9, 6, 25, 35]}},
    ],
    "stochastic_evolution": [
        {"src": "runmat-runtime/src/builtins/stats/random/stochastic_evolution.rs:40-51 cpu_fallback_handles_zero_scale (tol 1e-12): S*exp(drift*steps)",
         "state": [1.0, 2.0], "drift": 0.1, "scale": 0.0, "steps": 3},
    ],
    # runmat-runtime/src/builtins/math/linalg/solve/linsolve.rs tests (tolerance 1e-7, approx_eq :1111-1113)
    "linsolve": [
        {"src": "linsolve.rs:1209-1220 linsolve_basic_square", "a": {"shape": [2, 2], "data": [2, 1, 1, 2]}, "b": {"shape": [2, 1], "data": [4, 5]},
         "opts": {}, "out": {"shape": [2, 1], "data": [1, 2]}},
        {"src": "linsolve.rs:1224-1246 linsolve_lower_triangular_hint", "a": {"shape": [3, 3], "data": [3, -1, 4, 0, 2, 1, 0, 0, 5]},
         "b": {"shape": [3, 1], "data": [9, 1, 19]}, "opts": {"lower": True}, "out": {"shape": [3, 1], "data": [3, 2, 1]}},
        # linsolve_transposed_triangular_hint (:1250-1282) asserts equality with the general solve of A' x = b; A' = [3 1 0; 0 4 2; 0 0 5],
        # b = [5 14 23]'  ->  x3 = 4.6, x2 = (14 - 9.2)/4 = 1.2, x1 = (5 - 1.2)/3
        {"src": "linsolve.rs:1250-1282 linsolve_transposed_triangular_hint", "a": {"shape": [3, 3], "data": [3, 1, 0, 0, 4, 2, 0, 0, 5]},
         "b": {"shape": [3, 1], "data": [5, 14, 23]}, "opts": {"lower": True, "transposed": True},
         "out": {"shape": [3, 1], "data": [(5 - 1.2) / 3, 1.2, 4.6]}},
        {"src": "linsolve.rs:1434-1465 linsolve_recovers_rcond_output (identity, LT hint added: rcond = 1)", "a": {"shape": [2, 2], "data": [1, 0, 0, 1]},
         "b": {"shape": [2, 1], "data": [1, 2]}, "opts": {"lower": True, "need_rcond": True}, "out": {"shape": [2, 1], "data": [1, 2]}, "rcond": 1.0},
    ],
    "linspace": [
        {"src": "runmat-accelerate/src/simple_provider.rs:3488-3512 (last element forced to stop)", "start": 0.0, "stop": 1.0, "count": 5,
         "out": [0.0, 0.25, 0.5, 0.75, 1.0]},
    ],
    "rng": {
        "src": "runmat-runtime/src/builtins/common/random.rs:9-13 constants; DEFAULT_RNG_SEED 0x9e3779b97f4a7c15; stream values are pinned by "
               "algorithm only (random.rs:609-644 helpers are self-referential)",
        "default_seed": 0x9e3779b97f4a7c15, "multiplier": 6364136223846793005, "increment": 1, "shift": 11,
    },
}
out = Path(__file__).with_name("reference_kats.json")
out.write_text(json.dumps(kats, indent=1))
print("wrote", out)
