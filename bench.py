#!/usr/bin/env python
"""bench.py — measures the hot path on N B200s of one node (contract: see DESIGN.md "Measurement").

  python bench.py --gpus N --steps K --warmup W            our CUDA provider (through the C ABI)
  python bench.py --impl reference --steps K --warmup W    the reference's CPU path (oracle port), host cores

Workload (BASELINE.json configs[1]): fused chain + sum() reduction on 4096x4096 f64, per GPU.
One "step" = one pass of the hot path over one batch:
    (i)  C = sin(A) .* B + 1        fused elementwise, C materialised      24 B/elem algorithmic
    (ii) s = sum(sin(A).*B + 1)     fused one-pass reduction               16 B/elem algorithmic
At N > 1 each rank owns an independent batch (weak scaling, no data-path collective) and the per-rank sums
meet in ONE NCCL all-reduce of a single f64 per step (the path's only exchange).
`value` = algorithmic bytes of all ranks / device time (CUDA events, max over ranks): achieved HBM GB/s.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

N_SIDE = 4096
ELEMS = N_SIDE * N_SIDE
BYTES_EW = 24 * ELEMS   # read A, read B, write C
BYTES_RED = 16 * ELEMS  # read A, read B (+8 B result)
BYTES_STEP = BYTES_EW + BYTES_RED
METRIC = "achieved HBM GB/s, fused sin(A).*B+1 (+sum) on 4096x4096 f64 per GPU"


def measured_peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        d = json.loads(f.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples SM clocks / clock-event reasons DURING the timed region (B200_PROFILING.md recipe).

    The timed region of this workload is a few milliseconds, far shorter than one `nvidia-smi -lms` period, so the primary
    sampler is an NVML polling thread (nvidia_ml_py, ~2 kHz; the main thread sits in GIL-free ctypes calls while the GPU works).
    `nvidia-smi -lms 100` runs beside it as the fallback / cross-check."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None
        self.thread = None
        self.stop_flag = threading.Event()
        self.nvml_sm, self.nvml_reasons, self.nvml_max = [], 0, None
        self.source = None

    def _nvml_handle(self):
        import pynvml

        pynvml.nvmlInit()
        try:  # CUDA ordinal -> NVML handle by PCI bus id (robust under CUDA_VISIBLE_DEVICES)
            import torch

            pr = torch.cuda.get_device_properties(self.gpu)
            bus = f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
            return pynvml, pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        except Exception:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            idx = self.gpu
            if vis and all(x.strip().isdigit() for x in vis.split(",")):
                idx = int(vis.split(",")[self.gpu])
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(idx)

    def _poll(self, pynvml, h):
        get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag.is_set():
            try:
                self.nvml_sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                self.nvml_reasons |= int(get_reasons(h))
            except Exception:
                break
            time.sleep(0.0005)

    def start(self):
        try:
            if os.environ.get("RUNMAT_B200_NO_CLOCK_THREAD"):
                raise RuntimeError("NVML polling disabled")
            pynvml, h = self._nvml_handle()
            self.nvml_max = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self._pynvml = pynvml
            self.thread = threading.Thread(target=self._poll, args=(pynvml, h), daemon=True)
            self.thread.start()
        except Exception:
            self.thread = None
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.thread:
            self.stop_flag.set()
            self.thread.join(timeout=2)
        sm, mx, reasons = [], [], set()
        if self.proc:
            if not self.nvml_sm:
                time.sleep(0.15)  # no NVML samples: give nvidia-smi the chance to emit at least one line
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
            for line in Path(self.path).read_text().splitlines():
                parts = [x.strip() for x in line.split(",")]
                if len(parts) < 9:
                    continue
                try:
                    sm.append(float(parts[1]))
                    mx.append(float(parts[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        if self.nvml_sm:
            pn = self._pynvml
            for name, attr in (("hw_slowdown", "nvmlClocksThrottleReasonHwSlowdown"), ("hw_thermal_slowdown", "nvmlClocksThrottleReasonHwThermalSlowdown"),
                               ("sw_thermal_slowdown", "nvmlClocksThrottleReasonSwThermalSlowdown"), ("sw_power_cap", "nvmlClocksThrottleReasonSwPowerCap")):
                if self.nvml_reasons & int(getattr(pn, attr, 0)):
                    reasons.add(name)
            return {"sm_mhz": statistics.median(self.nvml_sm), "sm_max_mhz": self.nvml_max or (max(mx) if mx else None), "reasons": sorted(reasons),
                    "samples": len(self.nvml_sm), "source": "nvml", "nvidia_smi_samples": len(sm)}
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock samples (nvml and nvidia-smi unavailable)"], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


WORKLOAD = "elementwise-math fused chain + sum(), 4096x4096 f64 per GPU (BASELINE.json configs[1])"
# identical in both arms (the driver compares the dicts); per-arm details live in the sibling key "run"
CONFIG = {"workload": WORKLOAD, "bytes_per_step_per_gpu": BYTES_STEP,
          "l2": "inputs (2x128 MiB) + output (128 MiB) per step exceed the 126 MB L2; no flush needed between iterations"}


def synth_inputs(rank: int):
    """Synthetic data of the named shape: A ~ U(0,4pi), B ~ U(-1,1), seeded per rank."""
    rng = np.random.default_rng(1234 + rank)
    A = rng.uniform(0.0, 4.0 * math.pi, ELEMS)
    B = rng.uniform(-1.0, 1.0, ELEMS)
    return A, B


# =====================================================================================================================
# reference arm: the reference's CPU implementation of the path (oracle port; the Rust reference cannot be built here)
# =====================================================================================================================
def cpu_reference_step(orc, A2, B2, one):
    """The unfused builtin sequence the reference CPU path executes (one host tensor per op, BroadcastPlan index math):
    sin (sin.rs:265-270) -> times (times.rs:682-700) -> plus -> sum 'all' (sum.rs:996-1079)."""
    t0 = orc.unary("sin", A2)
    t1 = orc.elem_binary("mul", t0, B2)
    C = orc.elem_binary("add", t1, one)
    s = orc.sum_dims(C, [0, 1])
    return C, float(s[0, 0])


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle_binding import Oracle

    orc = Oracle()
    A, B = synth_inputs(0)
    A2, B2 = A.reshape((N_SIDE, N_SIDE), order="F"), B.reshape((N_SIDE, N_SIDE), order="F")
    one = np.array([[1.0]])
    for _ in range(max(args.warmup, 0) and 1):
        cpu_reference_step(orc, A2, B2, one)
    times = []
    t_all = time.perf_counter()
    for _ in range(args.steps):
        t0 = time.perf_counter()
        cpu_reference_step(orc, A2, B2, one)
        times.append(time.perf_counter() - t0)
    total = time.perf_counter() - t_all
    value = BYTES_STEP * args.steps / total / 1e9
    sample = f"{args.steps} full passes of the 4096x4096 f64 step on 1 core (the reference CPU path is single-threaded: no rayon/BLAS on it)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": CONFIG,
        "run": {"arm": "reference CPU path (oracle port of the unfused builtin sequence), host cores"},
        "cpu_baseline": {"value": value, "unit": "GB/s", "cores": 1, "kind": "port", "sample": sample, "host_cores": os.cpu_count()},
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# =====================================================================================================================
# our arm
# =====================================================================================================================
class Exchange:
    """The sharded paths' only exchange: the sum of one scalar per rank. Preference order:
    'p2p'   peer-memory slots over NVLink (rm_comm_p2p_*: publish fused into the producing kernel, lazy combine),
    'nccl'  the provider's own NCCL communicator (rm_comm_allreduce_sum),
    'torch' torch.distributed on a side stream.
    Every rank takes the same branch (agreed with a MIN all-reduce)."""

    def __init__(self, p, rank, world, local_rank, dist, torch, label=""):
        self.p, self.rank, self.world, self.local_rank, self.dist, self.torch = p, rank, world, local_rank, dist, torch
        self.kind = "none"
        if world == 1:
            return
        dev = f"cuda:{local_rank}"

        def agree(ok):
            flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            return bool(flag.item())

        def probe():
            h = p.upload(np.array([float(rank + 1)]), (1, 1))
            got = float(p.download(p.comm_allreduce_sum(h))[0, 0])
            p.free(h)
            return got == world * (world + 1) / 2

        if not os.environ.get("RUNMAT_B200_NO_P2P") and not os.environ.get("RUNMAT_B200_TORCH_ALLREDUCE"):
            ok = True
            try:
                mine = p.comm_p2p_export()
                handles = [None] * world
                dist.all_gather_object(handles, mine)
                p.comm_p2p_connect(handles, rank, world)
                ok = probe() and probe()
            except Exception as exc:  # noqa: BLE001 - any failure selects the next exchange on ALL ranks
                print(f"[bench{label}] rank {rank}: peer-memory exchange unavailable ({exc})", file=sys.stderr)
                ok = False
            if agree(ok):
                self.kind = "p2p"
                return
        if not os.environ.get("RUNMAT_B200_TORCH_ALLREDUCE") and not p.comm_p2p_connected():
            ok = True
            try:
                from runmat_b200 import B200Provider

                ident = [B200Provider.comm_unique_id() if rank == 0 else None]
                dist.broadcast_object_list(ident, src=0)
                p.comm_init(ident[0], rank, world)
                ok = probe()
            except Exception as exc:  # noqa: BLE001
                print(f"[bench{label}] rank {rank}: native NCCL communicator unavailable ({exc}); using torch.distributed", file=sys.stderr)
                ok = False
            if agree(ok):
                self.kind = "nccl"
                return
        self.kind = "torch"
        self.prov_stream = torch.cuda.ExternalStream(p.stream(), device=dev)
        self.comm_stream = torch.cuda.current_stream()
        self.slots = torch.zeros(2, dtype=torch.float64 if p.precision() == "f64" else torch.float32, device=dev)
        self.copied = [torch.cuda.Event(), torch.cuda.Event()]
        self.reduced = [torch.cuda.Event(), torch.cuda.Event()]
        for e in self.reduced:
            e.record(self.comm_stream)
        self.n = 0

    def allreduce(self, h):
        """Global sum of the 1x1 handle `h` (consumed). Returns a handle that is valid once the exchange has landed
        (p2p / nccl: lazily ordered by its ready event; torch: the LOCAL handle, the global value sits in self.slots)."""
        if self.kind in ("p2p", "nccl"):
            g = self.p.comm_allreduce_sum(h)
            self.p.free(h)
            return g
        if self.kind == "torch":
            slot = self.n & 1
            self.n += 1
            self.prov_stream.wait_event(self.reduced[slot])
            self.p.copy_to_device(h, self.slots[slot:slot + 1].data_ptr(), 1)
            self.copied[slot].record(self.prov_stream)
            self.comm_stream.wait_event(self.copied[slot])
            self.dist.all_reduce(self.slots[slot:slot + 1])
            self.reduced[slot].record(self.comm_stream)
        return h

    def fence(self):
        """Order the provider's stream after every exchange issued so far (closes a timed region on the device)."""
        if self.kind in ("p2p", "nccl"):
            self.p.comm_fence()
        elif self.kind == "torch":
            self.prov_stream.wait_stream(self.comm_stream)

    def align(self, scratch):
        """Device-side rendezvous: one exchange whose result the provider's stream waits for, so every rank's timed region
        starts within microseconds of the others' (a host barrier leaves ~0.1-1 ms of skew, which a 15 ms Monte-Carlo shard
        then pays as waiting time inside its final exchange)."""
        if self.world > 1:
            g = self.allreduce(self.p.scalar_mul(scratch, 1.0))
            self.fence()
            self.p.free(g)

    def last_value(self, h):
        if self.kind == "torch":
            return float(self.slots[(self.n - 1) & 1].item())
        return float(self.p.download(h)[0, 0])

    def describe(self):
        return {"none": "single GPU", "p2p": "peer-memory slot exchange over NVLink (publish fused into the reduction's last block, rm_fused_reduction_allreduce)",
                "nccl": "provider NCCL communicator (rm_comm_allreduce_sum)", "torch": "torch.distributed all_reduce"}[self.kind]


def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    from runmat_b200 import B200Provider
    import fusion_text as ft
    from runmat_b200.provider import pinned_empty

    dist = None
    torch = None
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    p = B200Provider(local_rank, device_id=rank)
    from runmat_b200.sharding import bind_process_to_gpu_numa

    # before the host inputs / pinned buffers are first touched
    numa = None if os.environ.get("RUNMAT_B200_NO_NUMA_BIND") else bind_process_to_gpu_numa(p.pci_bus_id())
    ew_shader, red_shader = ft.sin_mul_add_wgsl(), ft.sum_sin_mul_add_wgsl()
    A, B = synth_inputs(rank)
    shape = (N_SIDE, N_SIDE)
    hA, hB = p.upload(A, shape), p.upload(B, shape)
    hOne = p.upload(np.array([1.0]), (1, 1))
    ex = Exchange(p, rank, world, local_rank, dist, torch)

    def step():
        hC = p.fused_elementwise(ew_shader, [hA, hB, hOne], shape, ELEMS)
        if ex.kind == "p2p":
            # ONE kernel: the per-rank reduction whose last block publishes the scalar into every peer's slot
            return hC, p.fused_reduction_allreduce(red_shader, [hA, hB], ELEMS)
        hS = p.fused_reduction(red_shader, [hA, hB], (1, 1), ELEMS, 1)
        return hC, ex.allreduce(hS)  # global sum of this step; overlaps the next step's kernels

    def sync_all():
        p.synchronize()
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        hC, hS = step()
        p.free(hC)
        p.free(hS)
    sync_all()

    # ---- timed region: exactly K steps, CUDA events on the launching stream, barrier+sync on both sides ----------
    t_before = p.telemetry_snapshot().kernel_launches
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sys.setswitchinterval(0.0005)  # let the NVML polling thread run between the (GIL-holding) launch calls
        sampler.start()
    sync_all()
    ex.align(hOne)
    p.timer_begin()
    in_flight = []  # N>1: a step's global sum is consumed (freed) SUM_DEPTH steps later, so its exchange has that much slack
    SUM_DEPTH = 4 if world > 1 else 1
    for _ in range(args.steps):
        hC, hS = step()
        p.free(hC)
        in_flight.append(hS)
        if len(in_flight) > SUM_DEPTH:
            p.free(in_flight.pop(0))
    last_sum = in_flight.pop()
    for h in in_flight:
        p.free(h)
    ex.fence()  # the timed region ends when the last exchange has landed
    ms = p.timer_end_ms()
    sync_all()
    clocks = sampler.stop() if rank == 0 else None
    launches = p.telemetry_snapshot().kernel_launches - t_before
    checksum = ex.last_value(last_sum) if world > 1 else float(p.download(last_sum)[0, 0])
    p.free(last_sum)
    hL = p.fused_reduction(red_shader, [hA, hB], (1, 1), ELEMS, 1)
    checksum_local = float(p.download(hL)[0, 0])
    p.free(hL)
    exchange_ok = None
    if world > 1:
        tt = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
        # the exchanged sum must equal the sum of the ranks' local values (folded here by NCCL in some order: 1e-12)
        tl = torch.tensor([checksum_local], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(tl)
        exchange_ok = bool(abs(float(tl.item()) - checksum) <= 1e-12 * abs(checksum)) and (ex.kind != "p2p" or p.comm_p2p_error() == 0)
    value = BYTES_STEP * world * args.steps / (ms * 1e-3) / 1e9

    # ---- roofline of the dominant kernel (fused elementwise, 24 B/elem): isolated launches, CUDA events ------------
    def time_kernel(fn, reps):
        def free(r):
            for h in (r if isinstance(r, (list, tuple)) else [r]):
                p.free(h)
        for _ in range(3):
            free(fn())
        p.synchronize()
        p.timer_begin()
        for _ in range(reps):
            free(fn())  # stream-ordered free: the pool hands the same block back, as in the step loop
        return p.timer_end_ms() / reps

    reps = max(args.steps, 10)
    # `roofline` times each kernel with PLAIN launches (what an ncu capture sees: one launch after the other). The step loop above and
    # `pipelined` below use the provider's default, programmatic dependent launch: the next kernel's CTAs become resident while
    # the previous kernel drains, so back-to-back launches overlap and the per-launch average is shorter than any single launch.
    p.set_launch_overlap(False)
    ew_ms = time_kernel(lambda: p.fused_elementwise(ew_shader, [hA, hB, hOne], shape, ELEMS), reps)
    red_ms = time_kernel(lambda: p.fused_reduction(red_shader, [hA, hB], (1, 1), ELEMS, 1), reps)
    p.set_launch_overlap(True)
    ew_pdl_ms = time_kernel(lambda: p.fused_elementwise(ew_shader, [hA, hB, hOne], shape, ELEMS), reps)
    red_pdl_ms = time_kernel(lambda: p.fused_reduction(red_shader, [hA, hB], (1, 1), ELEMS, 1), reps)

    def planner_pair():  # what the reference planner emits for sum(sin(A).*B+1): sin does not fold into a reduction (fusion.rs:1198-1258)
        hC = p.fused_elementwise(ew_shader, [hA, hB, hOne], shape, ELEMS)
        return [hC, p.reduce_sum(hC)]

    pair_ms = time_kernel(planner_pair, reps)
    peak, peak_src = measured_peaks()
    achieved = BYTES_EW / (ew_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "rm_fused_ew (C = sin(A).*B+1, 24 B/elem)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "traffic_source": None, "peak_source": peak_src, "us_per_launch": ew_ms * 1e3,
                "algorithmic_bytes_per_launch": BYTES_EW,
                "reduction_kernel": {"kernel": "rm_fused_red (sum(sin(A).*B+1), 16 B/elem)", "achieved": BYTES_RED / (red_ms * 1e-3) / 1e9,
                                     "frac": BYTES_RED / (red_ms * 1e-3) / 1e9 / peak, "us_per_launch": red_ms * 1e3},
                "pipelined": {"what": "the same launches back to back with programmatic dependent launch (provider default; the step loop runs this way): "
                                      "average time per launch, launches overlap head to tail",
                              "ew_us": ew_pdl_ms * 1e3, "ew_gbs": BYTES_EW / (ew_pdl_ms * 1e-3) / 1e9, "red_us": red_pdl_ms * 1e3,
                              "red_gbs": BYTES_RED / (red_pdl_ms * 1e-3) / 1e9},
                "planner_shaped_pair": {"what": "fused_elementwise (writes C) + reduce_sum(C): the two calls the reference planner emits for sum(sin(A).*B+1) "
                                                "(sin does not fold into a reduction, fusion.rs:1198-1258); 32 B/elem",
                                        "us": pair_ms * 1e3, "achieved": 32 * ELEMS / (pair_ms * 1e-3) / 1e9, "frac": 32 * ELEMS / (pair_ms * 1e-3) / 1e9 / peak,
                                        "vs_fused_16B_form_us": red_ms * 1e3}}
    profs = sorted((ROOT / "profiles").glob("r*_traffic.json"))
    prof = profs[-1] if profs else ROOT / "profiles" / "none"
    if prof.exists():
        try:
            roofline["traffic"] = json.loads(prof.read_text()).get("rm_fused_ew_dram_bytes_per_launch")
            roofline["traffic_source"] = f"committed ncu capture profiles/{prof.name} (dram__bytes_read.sum + dram__bytes_write.sum, per launch); not re-measured in this run"
        except Exception:
            pass

    # ---- e2e: the same step through the C ABI with HOST buffers (pinned), H2D + D2H inside the timed region ----------
    hostA, hostB, hostC = pinned_empty(ELEMS), pinned_empty(ELEMS), pinned_empty(ELEMS)
    hostA[:] = A
    hostB[:] = B

    # r06 probe (scripts/pcie_dev): the link gives ~53 GB/s H2D with a concurrent D2H, so a step cannot beat 268 MB / 53 GB/s = 5.06 ms
    # plus the un-overlapped first upload and last download. Uniform 8 chunks left 0.67 + 0.33 ms of that (5.9 ms/step); 32 uniform
    # chunks were host-bound (7.0 ms, r07). The schedule below tapers: small chunks at both ends (short pipeline fill and drain),
    # 1/8-size chunks in the middle (few calls). The per-chunk partial sums go to pinned host memory with async copies: one
    # synchronize per step.
    E2E_FRACS = [32, 32, 16, 8, 8, 8, 8, 8, 8, 16, 32, 32]  # chunk = ELEMS / f; sum of 1/f == 1 (r52 probe: 5.64 ms vs 5.72 for the 14-chunk 1/64 taper, 5.82 uniform 8)
    assert abs(sum(1.0 / f for f in E2E_FRACS) - 1.0) < 1e-12
    chunks, off = [], 0
    for f in E2E_FRACS:
        chunks.append((off, ELEMS // f))
        off += ELEMS // f
    assert off == ELEMS
    hostS = pinned_empty(len(chunks))

    def e2e_step():
        """Same step from HOST buffers through the C ABI, chunked and software-pipelined: the upload of chunk i+1 (H2D stream)
        is enqueued before the download of chunk i (compute stream), so H2D, kernels and D2H overlap on the full-duplex link."""
        pending = None
        for i in range(len(chunks) + 1):
            nxt = None
            if i < len(chunks):
                o, n = chunks[i]
                a = p.upload_ptr(hostA.ctypes.data + o * 8, (n, 1))
                b = p.upload_ptr(hostB.ctypes.data + o * 8, (n, 1))
                nxt = (i, a, b)
            if pending is not None:
                j, pa, pb, pc, ps = pending
                p.download_async_into_ptr(pc, hostC.ctypes.data + chunks[j][0] * 8, chunks[j][1])
                p.download_async_into_ptr(ps, hostS.ctypes.data + j * 8, 1)
                for h in (pa, pb, pc, ps):
                    p.free(h)
            if nxt is not None:
                i_, a, b = nxt
                n = chunks[i_][1]
                c = p.fused_elementwise(ew_shader, [a, b, hOne], (n, 1), n)
                s_ = p.fused_reduction(red_shader, [a, b], (1, 1), n, 1)
                pending = (i_, a, b, c, s_)
            else:
                pending = None
        p.synchronize()                      # C and the chunk sums are now complete in host memory
        val = 0.0
        for j in range(len(chunks)):
            val += float(hostS[j])
        if world > 1:
            tv = torch.tensor([val], dtype=torch.float64, device=f"cuda:{local_rank}")
            dist.all_reduce(tv)
            val = float(tv.item())
        return val

    e2e_step()
    sync_all()
    # Every e2e step ends with a stream synchronize + host reads, so per-step wall time is exact. The host is shared with other
    # tenants (PCIe/DRAM contention shows up as rare 2-3x outlier steps), so the reported figure is the MEDIAN step; the mean
    # over the same steps is kept beside it.
    e2e_steps = max(5, min(args.steps, 15))
    per_step = []
    e2e_val = 0.0
    for _ in range(e2e_steps):
        sync_all()
        t0 = time.perf_counter()
        e2e_val = e2e_step()
        per_step.append((time.perf_counter() - t0) * 1e3)
    sync_all()
    e2e_med = statistics.median(per_step)
    e2e_mean = sum(per_step) / len(per_step)
    if world > 1:
        tt = torch.tensor([e2e_med, e2e_mean], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_med, e2e_mean = float(tt[0].item()), float(tt[1].item())
    e2e_ms = e2e_med * e2e_steps
    e2e = {"value": BYTES_STEP * world * e2e_steps / (e2e_ms * 1e-3) / 1e9, "unit": "GB/s", "h2d_bytes_per_step": 2 * ELEMS * 8,
           "d2h_bytes_per_step": ELEMS * 8 + 8 * len(chunks), "ms_per_step": e2e_ms / e2e_steps, "ms_per_step_mean": e2e_mean, "ms_per_step_all": [round(x, 3) for x in per_step],
           "steps": e2e_steps, "statistic": "median over per-step host wall times (each step ends in a stream synchronize)",
           "note": "A,B uploaded from pinned host memory, fused elementwise + fused sum, C and the sum downloaded; 12 tapered chunks, software-pipelined (H2D stream / compute+D2H stream)",
           "checksum_rel_diff_vs_resident": (abs(e2e_val - checksum_local) / abs(checksum_local)) if world == 1 else None}

    # ---- host cost of one provider call (python ctypes + C dispatch, no synchronisation): says whether a step loop is host-bound
    def host_us(fn, reps=300):
        p.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            p.free(fn())
        dt = time.perf_counter() - t0
        p.synchronize()
        return dt / reps * 1e6

    tiny = p.upload(np.array([1.0, 2.0, 3.0, 4.0]), (4, 1))
    host_call = {"scalar_op_plus_free_us": host_us(lambda: p.scalar_mul(tiny, 2.0)),
                 "fused_elementwise_plus_free_us": host_us(lambda: p.fused_elementwise(ew_shader, [tiny, tiny, hOne], (4, 1), 4)),
                 "fused_reduction_plus_free_us": host_us(lambda: p.fused_reduction(red_shader, [tiny, tiny], (1, 1), 4, 1))}
    if ex.kind == "p2p":
        host_call["fused_reduction_allreduce_plus_free_us"] = host_us(lambda: p.fused_reduction_allreduce(red_shader, [tiny, tiny], 4))
        ex.fence()
    p.free(tiny)

    # ---- other configs of BASELINE.json, reported beside the headline (not the metric) --------------------------------
    extra = {}
    if not args.no_extra:
        try:
            extra = extra_workloads(p, ex, rank, world, local_rank, dist, torch)
        except Exception as e:  # an optional workload must never take the headline line down
            extra = {"error": f"{type(e).__name__}: {e}"}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle_binding import Oracle

        orc = Oracle()
        A2, B2 = A.reshape(shape, order="F"), B.reshape(shape, order="F")
        one = np.array([[1.0]])
        ts = []
        cpu_reference_step(orc, A2, B2, one)  # warm-up pass (page faults of the first large allocations)
        for _ in range(7):
            t0 = time.perf_counter()
            _, cpu_sum = cpu_reference_step(orc, A2, B2, one)
            ts.append(time.perf_counter() - t0)
        med = statistics.median(ts)
        cpu_baseline = {"value": BYTES_STEP / med / 1e9, "unit": "GB/s", "cores": 1, "kind": "port", "host_cores": os.cpu_count(),
                        "sample": "7 full passes (+1 warm-up) of the same 4096x4096 f64 step (median; ~6-15 s of CPU work), oracle port of the unfused CPU builtins, 1 core",
                        "seconds_per_step": med, "checksum_rel_diff_vs_gpu": abs(cpu_sum - checksum) / abs(cpu_sum)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": CONFIG,
            "run": {"arm": "B200 provider through the C ABI (ctypes)",
                    "parallelism": f"{world} independent batches, the per-rank sums meet in one scalar exchange per step: {ex.describe()}" if world > 1 else "single GPU",
                    "exchange": ex.kind, "exchange_verified": exchange_ok, "host_numa_binding": numa, "host_call_cost": host_call,
                    "timed_region": "device-side rendezvous, K steps, closing fence on the last exchange; CUDA events on the provider stream, max over ranks"},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "checksum": checksum, "extra": extra,
        }
        print(json.dumps(line))
    p.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def fp64_peak_tflops():
    """148 SMs x 64 FP64 FMA lanes x 2 flop x 1.965 GHz (no measured FP64 peak is provided; stated, not measured)."""
    return 148 * 64 * 2 * 1.965e9 / 1e12


def extra_workloads(p, ex, rank, world, local_rank, dist, torch):
    """The other BASELINE.json configs, each with its own roofline entry and stated work model:
    configs[4] Monte-Carlo 1e8 x 256 (sharded, one exchange), configs[3] 4K image batch B=64 (sharded by images, one exchange per
    batch), and at N=1: configs[2] matmul 8192^3, image_normalize / imfilter on the per-GPU share, mldivide."""
    out = {}
    peak_hbm, _ = measured_peaks()
    dev = f"cuda:{local_rank}"

    def max_over_ranks(ms):
        if world > 1:
            tt = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return float(tt.item())
        return ms

    def host_sync():
        p.synchronize()
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier()

    # ---- Monte-Carlo: paths sharded in contiguous ranges; RNG addressing is global, result independent of `world` ----------
    M, T = 100_000_000, 256
    lo, hi = rank * M // world, (rank + 1) * M // world
    drift, scale = (0.05 - 0.5 * 0.2 ** 2) / 252.0, 0.2 * math.sqrt(1.0 / 252.0)
    hS0 = p.fill((hi - lo, 1), 100.0)
    hScr = p.fill((1, 1), 0.0)
    ms = price = None
    for it in range(3):
        p.set_rng_state(0)
        host_sync()
        ex.align(hScr)                       # device-side rendezvous: ranks start within microseconds
        p.timer_begin()
        hS = p.stochastic_evolution_sharded(hS0, drift, scale, T, lo, M)
        hP = p.payoff_partial_sum(hS, 100.0)
        hG = ex.allreduce(hP)                # the path's only exchange: 1 f64 (stream-ordered, no host wait)
        ex.fence()
        ms = p.timer_end_ms()
        price = (ex.last_value(hG) if world > 1 else p.read_scalar(hG, 0)) / M * math.exp(-0.05 * T / 252.0)
        p.free(hS)
        p.free(hG)
    ms = max_over_ranks(ms)
    p.free(hS0)
    p.free(hScr)
    # work model: one path-step = 1/2 Box-Muller pair (-2 log u1, sqrt, sincos 2 pi u2 for two paths) + the noise-sum DFMA + the LCG hops.
    # The kernel is ISSUE-bound: 64.4 instruction slots per path-step, of which 18.1 go to the FP64 pipe (2 issue cycles each).
    psps = M * T / (ms * 1e-3)
    issue_rate = 148 * 4 * 32 * 1.965e9 * world  # thread-instructions/s at one warp instruction per clock per SM sub-partition
    pipe_rate = 148 * 64 * 1.965e9 * world       # FP64 lane-instructions/s (64 lanes/SM/clk)
    out["monte_carlo"] = {"paths": M, "steps": T, "ms": ms, "path_steps_per_s": psps, "price": price, "scaling": "strong",
                          "exchange": ex.describe() if world > 1 else None,
                          "roofline": {"bound": "instruction issue (1 warp instruction/clk/SM sub-partition; HBM traffic is 16 B/path total)",
                                       "work_model": f"{MC_INSTR_PER_PATH_STEP} executed instructions per path-step ({MC_FP64_INSTR_PER_PATH_STEP} of them DFMA+DMUL+DADD), "
                                                     "measured with ncu (smsp__inst_executed, SASS opcode counters; profiles/r40_mc_counters.txt, re-measured in profiles/r57_summary.md); "
                                                     "peak = 148 SM x 4 sub-partitions x 32 lanes x 1.965 GHz per GPU",
                                       "achieved": psps * MC_INSTR_PER_PATH_STEP / 1e12, "peak": issue_rate / 1e12, "unit": "T thread-instr/s",
                                       "frac": psps * MC_INSTR_PER_PATH_STEP / issue_rate,
                                       "fp64_pipe_frac": psps * MC_FP64_INSTR_PER_PATH_STEP / pipe_rate}}

    # ---- configs[3]: 4K image batch, B = 64 images sharded by image (batch is the stride-1 axis: each rank owns its own
    #      [B/N, H, W] tensor), per batch: image_normalize -> squared error vs input -> ONE scalar exchange (MSE) ------------------
    out["image_batch_64x4k_f32"] = image_batch_leg(p, rank, world, local_rank, dist, torch, peak_hbm)

    if world == 1:
        out.update(single_gpu_kernels(p, rank, local_rank, peak_hbm))
    return out


# Instructions evolve_kernel<double, lean, sum-of-log-returns> executes per path-step, MEASURED with ncu on the 1e8 x 256 run
# (r40_mc.ncu-rep, extract in profiles/r40_mc_counters.txt): smsp__inst_executed.sum = 5.151e10 warp instructions / (2.56e10 / 32)
# warp path-steps = 64.4; DFMA + DMUL + DADD thread instructions (3322.5 + 663.2 + 0.9 per clk over 1.1606e8 cycles) = 4.63e11 /
# 2.56e10 path-steps = 18.1 (round-2 start: 33.6 with the atanh-series log on converted uniforms; library math of round 1: 48.5).
# Issue slots are the binding resource (smsp__issue_active 75 %), the FP64 pipe is 42 % busy.
# r48 build (46-register budget): the loop body shrank from 129 to 121 SASS instructions per pair-step with the same 36 FP64 instructions;
# re-measured in r57 (r57_mc_lu.ncu-rep): smsp__inst_executed.sum = 4.8757e10 warp instructions / 8e8 warp path-steps = 60.9.
MC_INSTR_PER_PATH_STEP = 60.9
MC_FP64_INSTR_PER_PATH_STEP = 18.1


def image_batch_leg(p, rank, world, local_rank, dist, torch, peak_hbm):
    from runmat_b200 import B200Provider, ImageNormalizeDescriptor
    import fusion_text as ft
    from runmat_b200.sharding import batch_slices_for_rank, lcg_image_shard

    Bt, H, W = 64, 2160, 3840
    b0, b1 = batch_slices_for_rank(Bt, rank, world)
    Bl = b1 - b0
    dev = f"cuda:{local_rank}"
    with B200Provider(local_rank, device_id=rank + 1000, precision="f32") as p32:
        ex32 = Exchange(p32, rank, world, local_rank, dist, torch, label=" image")
        t0 = time.perf_counter()
        imgs = lcg_image_shard(Bt, H, W, b0, Bl)          # runmat_lcg.m:59-79, generated directly in this rank's layout
        gen_s = time.perf_counter() - t0
        hI = p32.upload(imgs, (Bl, H, W))
        del imgs
        d = ImageNormalizeDescriptor(Bl, H, W, 1e-6, gain=1.0123, bias=-0.02, gamma=1.8)
        O, I, T0, T1 = 0, 1, 10, 11
        mse_shader = ft.reduction_wgsl([O, I], [ft.FusionOp("primitive", "Sub", [O, I], T0), ft.FusionOp("primitive", "ElemMul", [T0, T0], T1)], T1,
                                       axis=0, scalar_ty="f32")
        n = Bl * H * W
        hScr = p32.fill((1, 1), 0.0)

        def batch_step():
            hO = p32.image_normalize(hI, d)
            hE = p32.fused_reduction(mse_shader, [hO, hI], (1, 1), n, 1)
            p32.free(hO)
            return ex32.allreduce(hE) if world > 1 else hE

        for _ in range(2):
            p32.free(batch_step())
        p32.synchronize()
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier()
        ex32.align(hScr)
        K = 5
        p32.timer_begin()
        hG = None
        for _ in range(K):
            if hG is not None:
                p32.free(hG)
            hG = batch_step()
        ex32.fence()
        ms = p32.timer_end_ms() / K
        sq = ex32.last_value(hG) if world > 1 else float(p32.download(hG)[0, 0])
        p32.free(hG)
        p32.free(hI)
        p32.free(hScr)
        if world > 1:
            tt = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        px = Bt * H * W
        alg = 20 * px  # normalise: 2 reads + 1 write (12 B/px, the second read cannot come from L2 at this size); MSE: 2 reads (8 B/px)
        return {"batch": Bt, "images_per_rank": Bl, "ms_per_batch": ms, "images_per_s": Bt / (ms * 1e-3), "mse": sq / px, "scaling": "strong",
                "exchange": ex32.describe() if world > 1 else None, "host_generate_s_per_rank": gen_s,
                "roofline": {"bound": "hbm", "work_model": "20 B/px: image_normalize 12 B/px (moments read + normalise read + write) + fused squared-error reduction 8 B/px",
                             "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak_hbm * world, "unit": "GB/s", "frac": alg / (ms * 1e-3) / 1e9 / (peak_hbm * world)}}


def single_gpu_kernels(p, rank, local_rank, peak_hbm):
    out = {}
    from runmat_b200 import B200Provider, ImageNormalizeDescriptor

    def timed(pp, fn, reps, warm_ms=40.0):
        """Device time per call (CUDA events). The GPU idles while the host synthesises inputs, so each measurement first keeps
        the device busy for ~warm_ms (clocks back at boost) before the timed back-to-back launches."""
        pp.free(fn())
        pp.synchronize()
        t0 = time.perf_counter()
        while (time.perf_counter() - t0) * 1e3 < warm_ms:
            for _ in range(8):
                pp.free(fn())
            pp.synchronize()
        pp.timer_begin()
        for _ in range(reps):
            pp.free(fn())
        return pp.timer_end_ms() / reps

    with B200Provider(local_rank, device_id=rank + 2000, precision="f32") as p32:
        Bi, H, W = 8, 2160, 3840
        img = np.random.default_rng(3).random(Bi * H * W, dtype=np.float32)
        hI = p32.upload(img, (Bi, H, W))
        d = ImageNormalizeDescriptor(Bi, H, W, 1e-6, gain=1.0123, bias=-0.02, gamma=1.8)
        ms = timed(p32, lambda: p32.image_normalize(hI, d), 20)
        px = Bi * H * W
        out["image_normalize_8x4k_f32"] = {"ms": ms, "roofline": {"bound": "hbm", "work_model": "12 B/px (2 reads + 1 write; 8 B/px would need the batch to stay in L2)",
                                                                   "achieved": 12 * px / (ms * 1e-3) / 1e9, "peak": peak_hbm, "unit": "GB/s", "frac": 12 * px / (ms * 1e-3) / 1e9 / peak_hbm,
                                                                   "frac_at_8B_per_px": 8 * px / (ms * 1e-3) / 1e9 / peak_hbm}}
        ms = timed(p32, lambda: p32.reduce_mean_nd(hI, [1, 2]), 20)
        out["mean_dims23_8x4k_f32"] = {"ms": ms, "kernel": "rm_fused_red (Strided layout, 8 slices)",
                                       "roofline": {"bound": "hbm", "work_model": "4 B/px", "achieved": 4 * px / (ms * 1e-3) / 1e9, "peak": peak_hbm, "unit": "GB/s",
                                                    "frac": 4 * px / (ms * 1e-3) / 1e9 / peak_hbm}}
        p32.free(hI)
        frame = np.random.default_rng(4).random(H * W * 3, dtype=np.float32)
        g = np.exp(-((np.arange(5) - 2)[:, None] ** 2 + (np.arange(5) - 2)[None, :] ** 2) / 2.0)
        hF, hK = p32.upload(frame, (H, W, 3)), p32.upload((g / g.sum()).astype(np.float32))
        ms = timed(p32, lambda: p32.imfilter(hF, hK, padding="replicate"), 20)
        smp = H * W * 3
        out["imfilter_5x5_4k_rgb_f32"] = {"ms": ms, "roofline": {"bound": "hbm (8 B/sample) / FP32 issue (25 unfused mul + 25 add per sample, bit-exact with the host order)",
                                                                  "work_model": "8 B/sample", "achieved": 8 * smp / (ms * 1e-3) / 1e9, "peak": peak_hbm, "unit": "GB/s",
                                                                  "frac": 8 * smp / (ms * 1e-3) / 1e9 / peak_hbm}}
    n = 8192
    rng = np.random.default_rng(7)
    hA = p.upload(rng.uniform(-1, 1, n * n), (n, n))
    hB = p.upload(rng.uniform(-1, 1, n * n), (n, n))
    ms = timed(p, lambda: p.matmul(hA, hB), 3, warm_ms=0.0)
    st = p.ozaki_stats()
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    bf16 = float(peaks.get("bf16_tflops", 2250.0 * 0.76))
    ng = st.get("int8_gemms") or 29  # digit products + the accuracy guard's magnitude product, all exact int8 GEMMs (reported by the library)
    int8_ops = ng * 2.0 * n ** 3
    out["matmul_8192_f64"] = {"ms": ms, "gflops": 2.0 * n ** 3 / (ms * 1e-3) / 1e9, "fp64_fallback_tiles": st["fp64_tiles"], "int8_gemms": ng,
                              "engine": f"auto -> tcgen05 (Ozaki int8 split: {ng - 1} exact int8 digit GEMMs + 1 guard GEMM, TMEM int32 accumulate, f64 recombine; "
                                        "8-bit digits x 6 slices for K <= 21760, 7-bit x 7 beyond; device-side accuracy guard, no host sync)",
                              "roofline": {"bound": "tensor", "work_model": f"{ng} int8 GEMMs of 2*8192^3 op; peak = 2 x the measured bf16 burst (int8 rate = 2 x bf16 on tcgen05)",
                                           "achieved": int8_ops / (ms * 1e-3) / 1e12, "peak": 2 * bf16, "unit": "TOP/s", "frac": int8_ops / (ms * 1e-3) / 1e12 / (2 * bf16),
                                           "f64_equivalent_tflops": 2.0 * n ** 3 / (ms * 1e-3) / 1e12}}
    p.set_matmul_engine(1)
    ms1 = timed(p, lambda: p.matmul(hA, hB), 1, warm_ms=0.0)
    p.set_matmul_engine(0)
    out["matmul_8192_f64_dmma"] = {"ms": ms1, "gflops": 2.0 * n ** 3 / (ms1 * 1e-3) / 1e9, "engine": "FP64 DMMA mma.sync.m8n8k4",
                                   "roofline": {"bound": "fp64 tensor pipe", "work_model": "2*8192^3 flop", "achieved": 2.0 * n ** 3 / (ms1 * 1e-3) / 1e12, "peak": fp64_peak_tflops(),
                                                "unit": "TFLOP/s", "frac": 2.0 * n ** 3 / (ms1 * 1e-3) / 1e12 / fp64_peak_tflops(), "peak_source": "stated: 148 SM x 64 FMA/clk x 1.965 GHz"}}
    p.free(hA)
    p.free(hB)
    # mldivide: 4096 x 4096 system, 64 right-hand sides (device LU with partial pivoting)
    ns = 4096
    hM = p.upload(rng.uniform(-1, 1, ns * ns) + np.eye(ns).reshape(-1) * 4.0, (ns, ns))
    hR = p.upload(rng.uniform(-1, 1, ns * 64), (ns, 64))
    p.free(p.mldivide(hM, hR))
    p.synchronize()
    t0 = time.perf_counter()
    hX = p.mldivide(hM, hR)
    p.synchronize()
    ms = (time.perf_counter() - t0) * 1e3
    flop = (2.0 / 3.0) * ns ** 3 + 2.0 * ns * ns * 64
    out["mldivide_4096_x64"] = {"ms": ms, "gflops_lu": (2.0 / 3.0) * ns ** 3 / (ms * 1e-3) / 1e9,
                                "roofline": {"bound": "fp64 tensor pipe (updates) + latency (panel)", "work_model": "2/3 n^3 + 2 n^2 nrhs flop", "achieved": flop / (ms * 1e-3) / 1e12,
                                             "peak": fp64_peak_tflops(), "unit": "TFLOP/s", "frac": flop / (ms * 1e-3) / 1e12 / fp64_peak_tflops()}}
    for h in (hM, hR, hX):
        p.free(h)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.steps == 50:
            args.steps = 5
        return run_reference(args)
    return run_ours(args)


class _CleanStdout:
    """Everything libraries write to fd 1 while the bench runs (NCCL prints its version banner there) goes to stderr; the JSON line is
    written to the real stdout, so stdout carries exactly ONE line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        self.real = os.fdopen(self.saved, "w")
        self.py_stdout, sys.stdout = sys.stdout, self.real
        return self

    def __exit__(self, *exc):
        self.real.flush()
        sys.stdout = self.py_stdout
        os.dup2(self.saved, 1)
        return False


if __name__ == "__main__":
    with _CleanStdout():
        rc = main()
    sys.exit(rc)
