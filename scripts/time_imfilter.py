"""Times imfilter (BASELINE config[3]: 3840x2160x3 f32) on the TMA-staged persistent kernel, the per-tile packed kernel and the generic tiled kernel."""
import os, sys, numpy as np
sys.path.insert(0, '.')
from runmat_b200 import B200Provider
p = B200Provider(0, precision="f32")
rng = np.random.default_rng(0)
img = rng.uniform(0, 1, (2160, 3840, 3)).astype(np.float32)
hi = p.upload(img)
for K in (3, 5, 7):
    hk = p.upload(rng.uniform(0, 1, (K, K)).astype(np.float32))
    for name, env in (("tma-staged packed rbw8", {}), ("tma-staged packed rbw4", {"RUNMAT_B200_IMFILTER_RBW": "4"}), ("tma-staged mixed rbw4", {"RUNMAT_B200_IMFILTER_RBW": "4", "RUNMAT_B200_IMFILTER_MODE": "2"}),
                      ("tma-staged scalar rbw4", {"RUNMAT_B200_IMFILTER_RBW": "4", "RUNMAT_B200_IMFILTER_MODE": "0"}),
                      ("per-tile packed", {"RUNMAT_B200_IMFILTER_NO_TMA": "1"}), ("generic tiled", {"RUNMAT_B200_IMFILTER_GENERIC": "1"})):
        for k in ("RUNMAT_B200_IMFILTER_NO_TMA", "RUNMAT_B200_IMFILTER_GENERIC", "RUNMAT_B200_IMFILTER_RBW", "RUNMAT_B200_IMFILTER_CTAS", "RUNMAT_B200_IMFILTER_MODE"):
            os.environ.pop(k, None)
        os.environ.update(env)
        for _ in range(3): p.free(p.imfilter(hi, hk, padding="replicate"))
        p.flush_l2(); p.timer_begin()
        for _ in range(10): p.free(p.imfilter(hi, hk, padding="replicate"))
        ms = p.timer_end_ms() / 10
        px = img.size
        print(f"imfilter {K}x{K} {name}: {ms:.4f} ms  {px * 8 / ms / 1e6:.0f} GB/s (8 B/sample)")
print("device flags", p.device_flags())
