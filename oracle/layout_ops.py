"""TEST INFRASTRUCTURE ONLY — numpy restatement of the reference's host semantics for the bit-exact indexing / layout class
(SURVEY §8f row 2). Nothing in the product path imports this module; tests use it as the checker for `rm_find`,
`rm_sub2ind`, `rm_ind2sub`, `rm_permute`, `rm_repmat`, `rm_cat`, `rm_eye`.

Paths below are relative to /root/reference/crates/runmat-runtime/src/builtins. All tensors are column-major (order="F");
indices are MATLAB-style, 1-based, carried as f64 like the reference does. Pinned against the reference's own unit-test
literals in tests/golden/reference_kats.json (groups find / sub2ind / ind2sub / permute / repmat / cat / eye).
"""
from __future__ import annotations

import numpy as np


def _flat(a):
    return np.asarray(a, dtype=np.float64).reshape(-1, order="F")


def find(x, limit=None, direction="first"):
    """array/indexing/find.rs:593-629 (compute_find, real storage): walk the column-major data forwards ('first') or
    backwards ('last'), record idx+1 where value != 0.0 (NaN counts as non-zero), stop at `limit`.
    Returns (linear, rows, cols, values) as f64 column vectors; rows/cols follow find.rs' FindResult (2-D view: rows = dim 1,
    cols = remaining dims folded)."""
    x = np.asarray(x, dtype=np.float64)
    data = _flat(x)
    hits = []
    if limit != 0:
        order = range(data.size) if direction == "first" else range(data.size - 1, -1, -1)
        for i in order:
            if data[i] != 0.0:
                hits.append(i + 1)
                if limit is not None and len(hits) >= limit:
                    break
    lin = np.asarray(hits, dtype=np.float64)
    rows_n = x.shape[0] if x.ndim >= 1 and x.size else 1
    zero = (lin - 1).astype(np.int64)
    rows = (zero % max(rows_n, 1) + 1).astype(np.float64)
    cols = (zero // max(rows_n, 1) + 1).astype(np.float64)
    vals = data[zero] if lin.size else np.zeros(0)
    return lin, rows, cols, vals


class IndexError_(ValueError):
    pass


def _coerce_subscript(v, dim_number, dim_size):
    """array/indexing/sub2ind.rs:398-429."""
    msg = "Subscript indices must either be real positive integers or logicals."
    if not np.isfinite(v):
        raise IndexError_(msg)
    r = np.round(v)
    if abs(r - v) > np.finfo(np.float64).eps or r < 1.0:
        raise IndexError_(msg)
    if r > dim_size:
        names = {1: "rows", 2: "columns", 3: "pages"}
        raise IndexError_(f"Index exceeds the number of {names[dim_number]} in dimension {dim_number}." if dim_number in names
                          else "Index exceeds array dimensions.")
    return int(r)


def sub2ind(dims, subs):
    """array/indexing/sub2ind.rs:336-388 (compute_indices): scalars broadcast over the (single) non-scalar shape, column-major
    strides, 1-based result; all-scalar input gives a [1,1] result."""
    dims = [int(d) for d in dims]
    subs = [np.asarray(s, dtype=np.float64) for s in subs]
    if len(subs) != len(dims):
        raise IndexError_("The number of subscripts supplied must equal the number of dimensions in the size vector.")
    shape = None
    for s in subs:
        if s.size != 1:
            if shape is not None and s.shape != shape:
                raise IndexError_("Subscript inputs must have the same size.")
            shape = s.shape
    n = 1 if shape is None else int(np.prod(shape))
    if shape is None:
        shape = (1, 1)
    strides = [1]
    for d in dims[:-1]:
        strides.append(strides[-1] * d)
    flat = [_flat(s) for s in subs]
    out = np.empty(n, dtype=np.float64)
    for i in range(n):
        off = 0
        for k, (d, s) in enumerate(zip(dims, flat)):
            v = s[0] if s.size == 1 else s[i]
            off += (_coerce_subscript(v, k + 1, d) - 1) * strides[k]
        out[i] = off + 1
    return out.reshape(shape, order="F")


def ind2sub(dims, idx):
    """array/indexing/ind2sub.rs:289-360 (compute_subscripts + coerce_linear_index): one output per dimension, each shaped
    like `idx`; indices must be positive integers <= prod(dims)."""
    dims = [int(d) for d in dims]
    idx = np.asarray(idx, dtype=np.float64)
    total = int(np.prod(dims))
    flat = _flat(idx)
    outs = [np.empty(flat.size, dtype=np.float64) for _ in dims]
    for i, v in enumerate(flat):
        if not np.isfinite(v) or abs(np.round(v) - v) > np.finfo(np.float64).eps:
            raise IndexError_("Linear indices must be positive integers.")
        r = int(np.round(v))
        if r < 1:
            raise IndexError_("Linear indices must be positive integers.")
        if r > total:
            raise IndexError_("Index exceeds number of array elements. Index must not exceed %d." % total)
        rem = r - 1
        for k, d in enumerate(dims):
            outs[k][i] = rem % d + 1
            rem //= d
    shape = idx.shape if idx.ndim >= 2 else (1, 1) if idx.ndim == 0 else (idx.size, 1)
    return [o.reshape(shape, order="F") for o in outs]


def permute(x, order_one_based):
    """array/shape/permute.rs:330-355, :498-564 (permute_generic): order may name dims beyond the input rank (trailing singleton
    dims are added); out[..., i_order[k], ...] = in[...]."""
    x = np.asarray(x, dtype=np.float64)
    order = [int(o) - 1 for o in order_one_based]
    if sorted(order) != list(range(len(order))):
        raise ValueError("permute: order must be a permutation of 1:n")
    if len(order) < x.ndim:
        raise ValueError("permute: order length must be at least ndims(A)")
    shape = list(x.shape) + [1] * (len(order) - x.ndim)
    return np.transpose(x.reshape(shape, order="F"), order)


def repmat(x, reps):
    """array/shape/repmat.rs:228-304, :504-507: tile along each dimension; reps shorter than the rank are padded with 1, a
    single scalar rep n means [n, n] (repmat.rs:367-384)."""
    x = np.asarray(x, dtype=np.float64)
    reps = [int(r) for r in reps]
    if len(reps) == 1:
        reps = [reps[0], reps[0]]
    nd = max(x.ndim, len(reps))
    xs = x.reshape(list(x.shape) + [1] * (nd - x.ndim), order="F")
    return np.tile(xs, reps + [1] * (nd - len(reps)))


def cat(dim_one_based, arrays):
    """array/shape/cat.rs (numeric path, :560-760): all dims except `dim` must agree (after padding with trailing 1s)."""
    d = int(dim_one_based) - 1
    arrs = [np.asarray(a, dtype=np.float64) for a in arrays]
    nd = max(max(a.ndim for a in arrs), d + 1)
    arrs = [a.reshape(list(a.shape) + [1] * (nd - a.ndim), order="F") for a in arrs]
    ref = list(arrs[0].shape)
    for a in arrs[1:]:
        for k in range(nd):
            if k != d and a.shape[k] != ref[k]:
                raise ValueError("cat: dimension mismatch")
    return np.concatenate(arrs, axis=d)


def eye(rows, cols=None):
    """array/creation/eye.rs:330-420: ones on the main diagonal of a rows x cols matrix (column-major)."""
    cols = rows if cols is None else cols
    out = np.zeros((int(rows), int(cols)), dtype=np.float64, order="F")
    for i in range(min(int(rows), int(cols))):
        out[i, i] = 1.0
    return out
