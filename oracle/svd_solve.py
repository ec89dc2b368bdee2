"""CPU restatement of the reference's mldivide for real matrices (TEST INFRASTRUCTURE ONLY — see rm_oracle.cpp header).

Reference: crates/runmat-runtime/src/builtins/math/linalg/ops/mldivide.rs:317-404
  solve_real_matrix: SVD::new(lhs, true, true); tol = compute_svd_tolerance(sv, rows, cols); svd.solve(rhs, tol)
  compute_svd_tolerance = f64::EPSILON * max(rows, cols) * max(max_sv, 1.0)
The SVD itself lives in a third-party dependency that is absent from /root/reference: nalgebra 0.32.6 (Cargo.lock:3871-3872),
`linalg::SVD::solve(b, eps)`: x = V * diag(1/s_i if s_i > eps else 0) * U^T * b. Restated here with numpy's LAPACK SVD
(mathematically the same pseudo-inverse solve; singular vectors may differ by sign/rotation within equal singular values,
which cancels in the product). Parity is UNPINNED at the bit level: the reference's own tests pin mldivide only by residual
(mldivide.rs:662-676, ||A*X - B|| < 1e-12 on a 2x2 system; :680-696 least squares < 1e-10).
"""
import numpy as np


def mldivide(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.shape == (1, 1):  # mldivide.rs:321-325: rhs * (1/a)
        return b * (1.0 / a[0, 0])
    if a.shape[0] != b.shape[0]:
        raise ValueError("mldivide: row mismatch")
    if a.shape[0] == 0:
        return np.zeros((a.shape[1], b.shape[1]))
    u, s, vt = np.linalg.svd(a, full_matrices=False)
    tol = np.finfo(np.float64).eps * max(a.shape) * max(s.max() if s.size else 0.0, 1.0)
    inv = np.where(s > tol, 1.0 / np.where(s > tol, s, 1.0), 0.0)
    return vt.T @ (inv[:, None] * (u.T @ b))
