#!/usr/bin/env python
"""Summarises gpurun_out/<tag>_* ncu captures into profiles/ (tracked). Run on the CPU box: ncu -i needs no GPU."""
import collections
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
G, P = ROOT / "gpurun_out", ROOT / "profiles"
P.mkdir(exist_ok=True)

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed_pipe_fp64.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum", "lts__t_bytes.sum", "sm__cycles_elapsed.avg.per_second"]


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units = rows[hdr], rows[hdr + 1]
    return names, units, rows[hdr + 2:]


summary, traffic = [], {}
# ---- launch list ---------------------------------------------------------------------------------------------------
lf = G / f"{tag}_launches.csv"
if lf.exists():
    rows = list(csv.reader(open(lf)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) < 15:
            continue
        name = r[4].split("(")[0]
        agg.setdefault((name, r[8], r[7]), []).append(float(r[-1].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    with open(P / f"{tag}_launches.csv", "w") as f:
        f.write("kernel,grid,block,launches,avg_us,total_us,share_pct\n")
        for (name, grid, block), v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"\"{name}\",\"{grid}\",\"{block}\",{len(v)},{sum(v)/len(v)/1e3:.2f},{sum(v)/1e3:.1f},{sum(v)/tot*100:.2f}\n")
    summary.append(f"## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, bench.py --steps 4 --warmup 3)\n")
    summary.append("Cold-cache, serialised per-launch times: compare SHARES, not absolutes. Full table: `profiles/%s_launches.csv`.\n" % tag)
    summary.append("| kernel | grid | block | launches | avg µs | share |\n|---|---|---|---:|---:|---:|")
    for (name, grid, block), v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))[:14]:
        summary.append(f"| `{name[:70]}` | {grid} | {block} | {len(v)} | {sum(v)/len(v)/1e3:.1f} | {sum(v)/tot*100:.1f}% |")
    summary.append("")

# ---- full captures -------------------------------------------------------------------------------------------------
for rep in sorted(G.glob(f"{tag}_*.ncu-rep")):
    names, units, rows = raw_rows(rep)
    idx = {n: i for i, n in enumerate(names)}
    summary.append(f"## `{rep.name}` (`ncu --set full --clock-control none`)\n")
    for r in rows:
        kname = r[idx["Kernel Name"]].split("(")[0]
        summary.append(f"### {kname}  grid={r[idx['Grid Size']]} block={r[idx['Block Size']]}\n")
        summary.append("| metric | value | unit |\n|---|---:|---|")
        vals = {}
        for k in KEYS:
            if k in idx:
                vals[k] = r[idx[k]]
                summary.append(f"| {k} | {r[idx[k]]} | {units[idx[k]]} |")
        try:
            rd, wr = float(vals["dram__bytes_read.sum"].replace(",", "")), float(vals["dram__bytes_write.sum"].replace(",", ""))
            mul = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            rd *= mul.get(units[idx["dram__bytes_read.sum"]], 1.0)
            wr *= mul.get(units[idx["dram__bytes_write.sum"]], 1.0)
            dur = float(vals["gpu__time_duration.sum"].replace(",", ""))
            dur *= {"us": 1e-6, "ns": 1e-9, "ms": 1e-3, "s": 1.0}[units[idx["gpu__time_duration.sum"]]]
            summary.append(f"| **DRAM traffic (read+write)** | {rd + wr:.0f} | byte |")
            summary.append(f"| **DRAM GB/s under ncu** | {(rd + wr) / dur / 1e9:.0f} | GB/s |")
            traffic.setdefault(kname, []).append(rd + wr)
        except Exception as e:  # noqa
            pass
        summary.append("")

tj = {}
for k, v in traffic.items():
    tj[f"{k}_dram_bytes_per_launch"] = sum(v) / len(v)
(P / f"{tag}_traffic.json").write_text(json.dumps(tj, indent=1))
bench = G / f"{tag}_bench.json"
if bench.exists():
    (P / f"{tag}_bench.json").write_text(bench.read_text())
(P / f"{tag}_summary.md").write_text(f"# ncu summary {tag}\n\nGenerated by scripts/summarize_profiles.py from gpurun_out/{tag}_* (captured with scripts/profile.sh on a B200).\n\n" + "\n".join(summary) + "\n")
print((P / f"{tag}_summary.md").read_text()[:6000])
print(json.dumps(tj, indent=1))
