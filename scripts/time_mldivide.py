"""Times mldivide (n x n, 64 right-hand sides) with and without the two-stream look-ahead; residual check."""
import os, sys, time, numpy as np
sys.path.insert(0, '.')
from runmat_b200 import B200Provider
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
p = B200Provider(0)
rng = np.random.default_rng(0)
A = rng.uniform(-1, 1, (n, n)) + np.eye(n) * 4.0
B = rng.uniform(-1, 1, (n, 64))
hM, hR = p.upload(A), p.upload(B)
for name, env in (("look-ahead", {}), ("single stream", {"RUNMAT_B200_LU_NO_LOOKAHEAD": "1"})):
    os.environ.pop("RUNMAT_B200_LU_NO_LOOKAHEAD", None)
    os.environ.update(env)
    p.free(p.mldivide(hM, hR))
    p.synchronize()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); h = p.mldivide(hM, hR); p.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
        X = p.download(h); p.free(h)
    res = np.abs(A @ X - B).max() / (np.abs(A).max() * np.abs(X).max() * n)
    print(f"mldivide n={n} {name}: {min(ts):.2f} ms (median {sorted(ts)[2]:.2f}), scaled residual {res:.2e}")
