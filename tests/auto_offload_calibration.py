"""B200 threshold profile for the reference's native auto-offload.

TEST/MEASUREMENT TOOL (lives under tests/ because it times the CPU oracle). It repeats, outside Rust, the procedure of
`crates/runmat-accelerate/src/native_auto.rs:1627-1880` (`auto_calibrate`): for the reference's own probe sizes it
compares the CPU builtin (`plus`, `sum`, `mtimes` — here the oracle port, 1 core) with the provider call (`elem_add`,
`reduce_sum`, `matmul`) and picks the first size where the provider is faster. Two GPU clocks are reported: `issue`
(what the reference's `Instant::now()` around a ready-future call would see) and `sync` (call + stream synchronize).

It writes a calibration file in the format `apply_auto_offload_calibration_from_file` parses
(`native_auto.rs:330-416`: `auto_offload_calibration{runs,cpu_time_ms{..},units{..},provider{..}}`), plus the thresholds
the reference's procedure selects, so a maintainer can feed it to `runmat accel-calibrate --input <file>`.

    python tests/auto_offload_calibration.py [out.json]        # needs a B200
"""
from __future__ import annotations

import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from oracle_binding import Oracle  # noqa: E402
from runmat_b200 import B200Provider  # noqa: E402

ELEM_SIZES = [256, 1_024, 4_096, 16_384, 65_536]  # native_auto.rs:1651, :1724
MATMUL_DIMS = [32, 64, 96, 128, 192]  # native_auto.rs:1790


def best_of(fn, reps=7):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return min(ts)


def main() -> None:
    out_path = Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / "gpurun_out" / "b200_auto_offload_calibration.json"
    orc = Oracle()
    p = B200Provider(0)
    p.warmup()
    rows = {"elementwise": [], "reduction": [], "matmul": []}
    thresholds = {}

    def gpu_times(call):
        h = call()  # first call compiles / warms the kernel (the reference calls `warmup()` before calibrating)
        p.free(h)
        p.synchronize()
        issue, sync = [], []
        for _ in range(7):
            t0 = time.perf_counter()
            h = call()
            t1 = time.perf_counter()
            p.synchronize()
            t2 = time.perf_counter()
            p.free(h)
            issue.append(t1 - t0)
            sync.append(t2 - t0)
        return min(issue), min(sync)

    for n in ELEM_SIZES:
        a = np.arange(n, dtype=np.float64).reshape(n, 1)
        cpu = best_of(lambda: orc.elem_binary("add", a, a))
        ha, hb = p.upload(a), p.upload(a)
        issue, sync = gpu_times(lambda: p.elem_add(ha, hb))
        p.free(ha); p.free(hb)
        rows["elementwise"].append({"elements": n, "cpu_s": cpu, "gpu_issue_s": issue, "gpu_sync_s": sync})
    for n in ELEM_SIZES:
        a = np.arange(n, dtype=np.float64).reshape(n, 1)
        cpu = best_of(lambda: orc.reduce_sum(a))
        ha = p.upload(a)
        issue, sync = gpu_times(lambda: p.reduce_sum(ha))
        p.free(ha)
        rows["reduction"].append({"elements": n, "cpu_s": cpu, "gpu_issue_s": issue, "gpu_sync_s": sync})
    for n in MATMUL_DIMS:
        rng = np.random.default_rng(n)
        a, b = rng.uniform(-1, 1, (n, n)), rng.uniform(-1, 1, (n, n))
        cpu = best_of(lambda: orc.matmul(a, b, naive=True), reps=3)
        ha, hb = p.upload(a), p.upload(b)
        issue, sync = gpu_times(lambda: p.matmul(ha, hb))
        p.free(ha); p.free(hb)
        rows["matmul"].append({"n": n, "flops": n ** 3, "cpu_s": cpu, "gpu_issue_s": issue, "gpu_sync_s": sync})

    def first(rows_, key, size_key):
        for r in rows_:
            if r[key] < r["cpu_s"]:
                return r[size_key]
        return None  # usize::MAX in the reference: never offload on this criterion

    for clock in ("gpu_issue_s", "gpu_sync_s"):
        thresholds[clock] = {
            "binary_min_elems": first(rows["elementwise"], clock, "elements"),
            "reduction_min_elems": first(rows["reduction"], clock, "elements"),
            "matmul_min_flops": first(rows["matmul"], clock, "flops"),
        }

    # CPU cost coefficients from the largest probe of each family (what update_cpu_cost converges to)
    big_e, big_r, big_m = rows["elementwise"][-1], rows["reduction"][-1], rows["matmul"][-1]
    info = p.device_info_struct()
    doc = {
        "auto_offload_calibration": {
            "runs": 7,
            "cpu_time_ms": {"elementwise": big_e["cpu_s"] * 1e3, "reduction": big_r["cpu_s"] * 1e3, "matmul": big_m["cpu_s"] * 1e3},
            "units": {"elementwise": float(big_e["elements"]), "reduction": float(big_r["elements"]), "matmul_flops": float(big_m["flops"])},
            "provider": {"name": info.name.decode(), "vendor": info.vendor.decode(), "backend": info.backend.decode(), "device_id": int(info.device_id)},
        },
        "cpu_coefficients": {
            "cpu_elem_per_elem": big_e["cpu_s"] / big_e["elements"],
            "cpu_reduction_per_elem": big_r["cpu_s"] / big_r["elements"],
            "cpu_matmul_per_flop": big_m["cpu_s"] / big_m["flops"],
        },
        "thresholds_by_reference_procedure": thresholds,
        "probes": rows,
        "note": "CPU side = oracle port of the reference's single-threaded builtins on this box; GPU side = librm_accel_b200 through the C ABI.",
    }
    out_path.parent.mkdir(parents=True, exist_ok=True)
    out_path.write_text(json.dumps(doc, indent=1))
    print(json.dumps({"thresholds": thresholds, "cpu_coefficients": doc["cpu_coefficients"]}))


if __name__ == "__main__":
    main()
