"""Times the 8192^3 f64 matmul on the tcgen05 engine with 8-bit (6 slices) and 7-bit (7 slices) digits; SM clock sampled around each."""
import os, subprocess, sys, numpy as np
sys.path.insert(0, '.')
from runmat_b200 import B200Provider
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
p = B200Provider(0)
rng = np.random.default_rng(0)
hA, hB = p.upload(rng.uniform(-1, 1, n * n), (n, n)), p.upload(rng.uniform(-1, 1, n * n), (n, n))
def clk():
    return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap", "--format=csv,noheader", "-i", "0"], capture_output=True, text=True).stdout.strip()
for bits in ("8", "7", "8"):
    os.environ["RUNMAT_B200_OZAKI_BITS"] = bits
    for _ in range(2): p.free(p.matmul(hA, hB))
    p.synchronize(); p.timer_begin()
    for _ in range(5): p.free(p.matmul(hA, hB))
    mid = clk()
    ms = p.timer_end_ms() / 5
    st = p.ozaki_stats()
    print(f"bits {bits}: {ms:.3f} ms  {2.0 * n ** 3 / ms / 1e9:.1f} f64-equivalent TFLOP/s  int8 GEMMs {st['int8_gemms']}  {st['int8_gemms'] * 2.0 * n ** 3 / ms / 1e9:.0f} TOP/s   [{mid}]")
os.environ.pop("RUNMAT_B200_OZAKI_BITS", None)
for name, env in (("pair (cta_group::2)", {}), ("single CTA", {"RUNMAT_B200_OZAKI_1CTA": "1"}), ("pair (cta_group::2)", {})):
    os.environ.pop("RUNMAT_B200_OZAKI_1CTA", None)
    os.environ.update(env)
    for _ in range(2): p.free(p.matmul(hA, hB))
    p.synchronize(); p.timer_begin()
    for _ in range(5): p.free(p.matmul(hA, hB))
    mid = clk()
    ms = p.timer_end_ms() / 5
    st = p.ozaki_stats()
    print(f"{name}: {ms:.3f} ms  {2.0 * n ** 3 / ms / 1e9:.1f} f64-equivalent TFLOP/s  {st['int8_gemms'] * 2.0 * n ** 3 / ms / 1e9:.0f} TOP/s  err {st['pipeline_error']}  [{mid}]")
