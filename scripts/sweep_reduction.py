#!/usr/bin/env python
"""Sweeps the fused-reduction kernel's tuning knobs on the GPU box (each config in its own process: the knobs are read
when the kernel source is emitted). Prints µs/launch and GB/s for sum(sin(A).*B+1) on 4096x4096 f64 (268.4 MB)."""
import itertools
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, math, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r + '/tests')
from runmat_b200 import B200Provider
import fusion_text as ft
p = B200Provider(0)
n = 4096
rng = np.random.default_rng(0)
hA = p.upload(rng.uniform(0, 4*math.pi, n*n), (n, n)); hB = p.upload(rng.uniform(-1, 1, n*n), (n, n))
sh = ft.sum_sin_mul_add_wgsl()
plain = ft.reduction_wgsl([0], [], 0, axis=0)
muladd = ft.reduction_wgsl([0, 1], [ft.FusionOp("primitive", "ElemMul", [0, 1], 11), ft.FusionOp("primitive", "Add", [11, 2], 12)], 12, axis=0, const_values={2: 1.0})
for name, shader, ins, nbytes in (("sum(sin(A).*B+1)", sh, [hA, hB], 16*n*n), ("sum(A.*B+1)", muladd, [hA, hB], 16*n*n), ("sum(A)", plain, [hA], 8*n*n)):
    for _ in range(5): p.free(p.fused_reduction(shader, ins, (1, 1), n*n, 1))
    p.synchronize(); p.timer_begin()
    for _ in range(50): p.free(p.fused_reduction(shader, ins, (1, 1), n*n, 1))
    ms = p.timer_end_ms() / 50
    print(f"{name}: {ms*1e3:.1f} us {nbytes/ms/1e6:.0f} GB/s", end="  |  ")
h1 = p.upload(np.array([1.0]), (1, 1)); ew = ft.sin_mul_add_wgsl()
for _ in range(5): p.free(p.fused_elementwise(ew, [hA, hB, h1], (n, n), n*n))
p.synchronize(); p.timer_begin()
for _ in range(50): p.free(p.fused_elementwise(ew, [hA, hB, h1], (n, n), n*n))
ms = p.timer_end_ms() / 50
print(f"C=sin(A).*B+1: {ms*1e3:.1f} us {24*n*n/ms/1e6:.0f} GB/s")
''' % (ROOT, ROOT)
# (unroll, min CTAs/SM, CTAs per SM in the grid, libm trig, grid-stride split)
grid = [(2, 4, 4, 1, 0), (2, 4, 4, 0, 1), (2, 4, 4, 0, 0), (1, 4, 4, 0, 0), (2, 4, 8, 0, 0), (2, 5, 5, 0, 0), (2, 4, 4, 0, 1), (2, 4, 4, 0, 0)]
if len(sys.argv) > 1 and sys.argv[1] == "full":
    grid += [(u, m, b, 0, 0) for u, m, b in itertools.product((1, 2, 4), (4, 5, 6), (4, 8))]
for unroll, minb, bpsm, libm, gs in grid:
    env = dict(os.environ, RUNMAT_B200_RED_UNROLL=str(unroll), RUNMAT_B200_RED_MINB=str(minb), RUNMAT_B200_RED_BPSM=str(bpsm),
               RUNMAT_B200_LIBM_TRIG=str(libm), RUNMAT_B200_NO_KCACHE="1")
    if not gs: env["RUNMAT_B200_RED_BLOCKED"] = "1"
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print(f"libm_trig={libm} gridstride={gs} unroll={unroll} minb={minb} bpsm={bpsm}: {r.stdout.strip() or r.stderr.strip()[-300:]}", flush=True)
