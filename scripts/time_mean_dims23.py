"""mean(imgs,[2 3]) over [8,2160,3840] f32 (batch-fastest layout): times the Strided reduction kernel."""
import os, sys, numpy as np
sys.path.insert(0, '.')
from runmat_b200 import B200Provider
p = B200Provider(0, precision="f32")
x = np.random.default_rng(3).random(8 * 2160 * 3840, dtype=np.float32)
h = p.upload(x, (8, 2160, 3840))
for env in ("",):
    for _ in range(3): p.free(p.reduce_mean_nd(h, [1, 2]))
    p.synchronize(); p.timer_begin()
    for _ in range(20): p.free(p.reduce_mean_nd(h, [1, 2]))
    ms = p.timer_end_ms() / 20
    m = p.download(p.reduce_mean_nd(h, [1, 2])).reshape(-1)
    print(f"{ms*1e3:.1f} us  {x.nbytes/ms/1e6:.0f} GB/s", m[:3], float(np.abs(m - x.reshape(-1, 8).astype(np.float64).mean(axis=0)).max()))
