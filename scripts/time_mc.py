"""Times the 1e8-path x 256-step stochastic evolution (BASELINE configs[4]) with the lean math (mc_math.h) and with the CUDA math library."""
import os, sys, numpy as np
sys.path.insert(0, '.')
from runmat_b200 import B200Provider
p = B200Provider(0)
M, T = 100_000_000, 256
h = p.fill((M, 1), 100.0)
res = {}
for name, env in (("lean", {}), ("lean stepwise", {"RUNMAT_B200_MC_STEPWISE": "1"}), ("libm", {"RUNMAT_B200_MC_LIBM": "1"})):
    os.environ.pop("RUNMAT_B200_MC_LIBM", None); os.environ.pop("RUNMAT_B200_MC_STEPWISE", None)
    os.environ.update(env)
    p.set_rng_state(42)
    p.free(p.stochastic_evolution(h, 0.0001, 0.0126, 4))
    p.synchronize(); p.set_rng_state(42); p.timer_begin()
    out = p.stochastic_evolution(h, 0.0001, 0.0126, T)
    ms = p.timer_end_ms()
    x = p.download(out)[:200000, 0]
    res[name] = x
    print(f"{name}: {ms:.2f} ms  {M * T / ms / 1e6:.1f} G path-steps/s  mean {x.mean():.6f}")
    p.free(out)
for a in ("lean", "lean stepwise"):
    d = np.abs(res[a] - res["libm"]) / np.abs(res["libm"])
    print(f"{a} vs libm on the first 200000 paths: max rel diff {d.max():.3e}")
