"""One 5x5 imfilter on the 2160x3840x3 f32 frame, for an ncu capture (`ncu -k regex:imfilter -s 2 -c 1 ...`)."""
import sys, numpy as np
sys.path.insert(0, '.')
from runmat_b200 import B200Provider
p = B200Provider(0, precision="f32")
rng = np.random.default_rng(0)
hi = p.upload(rng.uniform(0, 1, (2160, 3840, 3)).astype(np.float32))
hk = p.upload(rng.uniform(0, 1, (5, 5)).astype(np.float32))
for _ in range(4):
    p.free(p.imfilter(hi, hk, padding="replicate"))
p.synchronize()
