"""Per-call device times of consecutive 8192^3 matmuls on the tcgen05 engine (pair kernel), with the engine's device flags."""
import os, sys, numpy as np
sys.path.insert(0, '.')
from runmat_b200 import B200Provider
n = 8192
p = B200Provider(0)
rng = np.random.default_rng(0)
hA, hB = p.upload(rng.uniform(-1, 1, n * n), (n, n)), p.upload(rng.uniform(-1, 1, n * n), (n, n))
for label, env in (("pair", {}), ("bits7 pair", {"RUNMAT_B200_OZAKI_BITS": "7"}), ("pair", {}), ("1cta", {"RUNMAT_B200_OZAKI_1CTA": "1"}), ("pair", {})):
    for k in ("RUNMAT_B200_OZAKI_BITS", "RUNMAT_B200_OZAKI_1CTA"): os.environ.pop(k, None)
    os.environ.update(env)
    ts = []
    for i in range(6):
        p.synchronize(); p.timer_begin()
        h = p.matmul(hA, hB)
        ms = p.timer_end_ms()
        st = p.ozaki_stats()
        ts.append(f"{ms:.1f}{'!' if st['pipeline_error'] else ''}")
        p.free(h)
    print(label, " ".join(ts))
