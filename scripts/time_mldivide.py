import sys, numpy as np
sys.path.insert(0, '.')
from runmat_b200 import B200Provider
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
p = B200Provider(0)
rng = np.random.default_rng(0)
hM = p.upload(rng.uniform(-1, 1, n * n) + np.eye(n).reshape(-1) * 4.0, (n, n))
hR = p.upload(rng.uniform(-1, 1, n * 64), (n, 64))
p.free(p.mldivide(hM, hR))
p.synchronize()
import time
t0 = time.perf_counter(); h = p.mldivide(hM, hR); p.synchronize(); print("mldivide ms", (time.perf_counter() - t0) * 1e3)
