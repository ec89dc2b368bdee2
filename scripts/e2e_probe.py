"""e2e leg of bench.py taken apart on the GPU box: what the link gives on THIS box (H2D alone, D2H alone, both) and what each chunk
schedule costs, through the same provider calls bench.py uses. Prints ms per step (median of 9)."""
import math, statistics, sys, time
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from runmat_b200 import B200Provider
from runmat_b200.provider import pinned_empty
import fusion_text as ft

N = 4096 * 4096
p = B200Provider(0)
rng = np.random.default_rng(0)
hostA, hostB, hostC = pinned_empty(N), pinned_empty(N), pinned_empty(N)
hostA[:] = rng.uniform(0, 4 * math.pi, N); hostB[:] = rng.uniform(-1, 1, N)
hOne = p.upload(np.array([[1.0]]))
ew, red = ft.sin_mul_add_wgsl(), ft.sum_sin_mul_add_wgsl()


def med(fn, reps=9):
    fn(); p.synchronize()
    ts = []
    for _ in range(reps):
        p.synchronize(); t0 = time.perf_counter(); fn(); p.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    return statistics.median(ts), min(ts)


def chunks_of(fracs):
    out, off = [], 0
    for f in fracs:
        out.append((off, N // f)); off += N // f
    assert off == N
    return out


def step(chunks, h2d=True, compute=True, d2h=True, keep=None):
    hostS = pinned_empty(len(chunks))
    def run():
        pending = None
        for i in range(len(chunks) + 1):
            nxt = None
            if i < len(chunks):
                o, n = chunks[i]
                if h2d:
                    a = p.upload_ptr(hostA.ctypes.data + o * 8, (n, 1)); b = p.upload_ptr(hostB.ctypes.data + o * 8, (n, 1))
                else:
                    a, b = keep[i]
                nxt = (i, a, b)
            if pending is not None:
                j, pa, pb, pc, ps = pending
                if d2h and pc is not None:
                    p.download_async_into_ptr(pc, hostC.ctypes.data + chunks[j][0] * 8, chunks[j][1])
                    p.download_async_into_ptr(ps, hostS.ctypes.data + j * 8, 1)
                for h in ((pa, pb) if h2d else ()) + ((pc, ps) if pc is not None else ()):
                    p.free(h)
            if nxt is not None:
                i_, a, b = nxt
                n = chunks[i_][1]
                if compute:
                    c = p.fused_elementwise(ew, [a, b, hOne], (n, 1), n); s_ = p.fused_reduction(red, [a, b], (1, 1), n, 1)
                else:
                    c = s_ = None
                pending = (i_, a, b, c, s_)
            else:
                pending = None
    return run


TAPER14 = [64, 64, 32, 16, 8, 8, 8, 8, 8, 8, 16, 32, 64, 64]
SCHEDULES = {"tapered 14 (bench.py)": TAPER14, "uniform 4": [4] * 4, "uniform 8": [8] * 8, "uniform 16": [16] * 16,
             "tapered 11 (quarters in the middle)": [64, 64, 32, 16, 4, 4, 4, 16, 32, 64, 64],
             "tapered 12 (1/32 ends)": [32, 32, 16, 8, 8, 8, 8, 8, 8, 16, 32, 32]}
for name, fr in list(SCHEDULES.items()):
    if abs(sum(1.0 / f for f in fr) - 1.0) > 1e-12:
        print(f"skip {name}: fractions sum to {sum(1.0 / f for f in fr)}"); del SCHEDULES[name]

print("link, whole arrays: H2D 268 MB  %.3f ms (min %.3f)" % med(step(chunks_of([1]), compute=False, d2h=False)))
ch = chunks_of(TAPER14)
print("tapered 14, H2D only            %.3f ms (min %.3f)" % med(step(ch, compute=False, d2h=False)))
print("tapered 14, H2D + kernels       %.3f ms (min %.3f)" % med(step(ch, d2h=False)))
keep = [(p.upload_ptr(hostA.ctypes.data + o * 8, (n, 1)), p.upload_ptr(hostB.ctypes.data + o * 8, (n, 1))) for o, n in ch]
print("tapered 14, kernels + D2H only  %.3f ms (min %.3f)" % med(step(ch, h2d=False, keep=keep)))
for a, b in keep: p.free(a); p.free(b)
for name, fr in SCHEDULES.items():
    print("full step, %-42s %.3f ms (min %.3f)" % ((name,) + med(step(chunks_of(fr)))))
