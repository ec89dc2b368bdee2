"""Times image_normalize on [8,2160,3840] and [64,2160,3840] f32 (clamp + gamma) with 256-bit and 128-bit accesses in the normalise sweep."""
import os, sys, numpy as np
sys.path.insert(0, '.')
from runmat_b200 import B200Provider, ImageNormalizeDescriptor
p = B200Provider(0, precision="f32")
rng = np.random.default_rng(0)
for B in (8, 64):
    h = p.upload(rng.uniform(0, 1, (B, 2160, 3840)).astype(np.float32))
    d = ImageNormalizeDescriptor(batch=B, height=2160, width=3840, epsilon=1e-6, gain=1.0123, bias=-0.02, gamma=1.8, clamp_zero=True)
    for name, env in (("256-bit end-to-front, evict-first moment loads", {}), ("256-bit end-to-front, plain moment loads", {"RUNMAT_B200_MOMENTS_KEEP": "1"}),
                      ("256-bit forward", {"RUNMAT_B200_NORMALIZE_FORWARD": "1"}), ("128-bit end-to-front", {"RUNMAT_B200_NORMALIZE_VEC4": "1"})):
        for k in ("RUNMAT_B200_NORMALIZE_VEC4", "RUNMAT_B200_NORMALIZE_FORWARD", "RUNMAT_B200_MOMENTS_KEEP"): os.environ.pop(k, None)
        os.environ.update(env)
        for _ in range(3): p.free(p.image_normalize(h, d))
        p.flush_l2(); p.synchronize(); p.timer_begin()
        for _ in range(10): p.free(p.image_normalize(h, d))
        ms = p.timer_end_ms() / 10
        print(f"image_normalize B={B} {name}: {ms:.4f} ms  {B * 2160 * 3840 * 12 / ms / 1e6:.0f} GB/s at 12 B/px")
    p.free(h)
