"""imfilter on 2160x3840x3 f32: register budget A/B for the 5x5 TMA kernel (2 vs 3 resident CTAs per SM) and the compute-warp phase
stagger (ns per co-resident CTA, ns for the upper warps)."""
import os, sys, numpy as np
sys.path.insert(0, '.')
from runmat_b200 import B200Provider
p = B200Provider(0, precision="f32")
rng = np.random.default_rng(0)
img = rng.uniform(0, 1, (2160, 3840, 3)).astype(np.float32)
hi = p.upload(img)
KNOBS = ("RUNMAT_B200_IMFILTER_STAGGER", "RUNMAT_B200_IMFILTER_WSTAGGER", "RUNMAT_B200_IMFILTER_MINB2", "RUNMAT_B200_IMFILTER_RBW")
for K in (5, 3, 7):
    hk = p.upload(rng.uniform(0, 1, (K, K)).astype(np.float32))
    for name, env in (("default", {}), ("91-register build (2 CTAs/SM)", {"RUNMAT_B200_IMFILTER_MINB2": "1"}), ("rbw4", {"RUNMAT_B200_IMFILTER_RBW": "4"}),
                      ("stagger 1000/0", {"RUNMAT_B200_IMFILTER_STAGGER": "1000"}), ("stagger 0/1000", {"RUNMAT_B200_IMFILTER_WSTAGGER": "1000"}),
                      ("stagger 1500/750", {"RUNMAT_B200_IMFILTER_STAGGER": "1500", "RUNMAT_B200_IMFILTER_WSTAGGER": "750"})):
        if K != 5 and "MINB2" in "".join(env): continue
        for k in KNOBS: os.environ.pop(k, None)
        os.environ.update(env)
        for _ in range(3): p.free(p.imfilter(hi, hk, padding="replicate"))
        p.flush_l2(); p.timer_begin()
        for _ in range(20): p.free(p.imfilter(hi, hk, padding="replicate"))
        ms = p.timer_end_ms() / 20
        print(f"imfilter {K}x{K} {name}: {ms * 1e3:.1f} us  {img.size * 8 / ms / 1e6:.0f} GB/s", flush=True)
print("device flags", p.device_flags())
