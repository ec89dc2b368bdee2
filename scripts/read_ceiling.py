"""Asymptotic read bandwidth of the fused reduction kernel: plain sum over 1 GiB (separates fixed launch/finish cost from streaming rate)."""
import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from runmat_b200 import B200Provider
import fusion_text as ft
p = B200Provider(0)
plain = ft.reduction_wgsl([0], [], 0, axis=0)
for logn in (24, 27, 29):
    n = 1 << logn
    h = p.fill((n, 1), 1.0)
    for _ in range(3): p.free(p.fused_reduction(plain, [h], (1, 1), n, 1))
    p.synchronize(); p.timer_begin()
    for _ in range(20): p.free(p.fused_reduction(plain, [h], (1, 1), n, 1))
    ms = p.timer_end_ms() / 20
    print(f"sum over 2^{logn} f64 ({8*n/2**20:.0f} MiB): {ms*1e3:.1f} us  {8*n/ms/1e6:.0f} GB/s")
    p.free(h)
