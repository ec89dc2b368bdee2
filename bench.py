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
        "data": "synthetic", "config": {"workload": WORKLOAD, "bytes_per_step_per_gpu": BYTES_STEP, "arm": "reference CPU path (oracle port of the unfused builtin sequence), host cores"},
        "cpu_baseline": {"value": value, "unit": "GB/s", "cores": 1, "kind": "port", "sample": sample, "host_cores": os.cpu_count()},
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# =====================================================================================================================
# our arm
# =====================================================================================================================
def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    from runmat_b200 import B200Provider, fusion_text as ft
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

    class CudaArray:  # zero-copy torch view of a provider buffer (for the NCCL all-reduce)
        def __init__(self, ptr, n):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 3}

    # Multi-GPU: the provider keeps its own stream; torch's current stream is the communication stream. Per step the 8-byte
    # sum is copied (one C call, stream-ordered) into a 2-slot persistent buffer, the comm stream waits on that copy and
    # all-reduces the slot in place, so the collective's latency overlaps the next step's kernels. A slot is reused two
    # steps later, after the provider stream has waited for its previous reduction.
    native_comm = False
    if world > 1 and not os.environ.get("RUNMAT_B200_TORCH_ALLREDUCE"):
        # Preferred exchange: the provider's own NCCL communicator (rm_comm_*): ONE C call per step issues the copy, the stream
        # dependencies and the all-reduce on the provider's communication stream. torch.distributed only ships the 128-byte id.
        ok = 1
        try:
            ident = [B200Provider.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(ident, src=0)
            p.comm_init(ident[0], rank, world)
            probe = p.upload(np.array([float(rank + 1)]), (1, 1))
            got = float(p.download(p.comm_allreduce_sum(probe))[0, 0])
            ok = 1 if got == world * (world + 1) / 2 else 0
        except Exception as exc:  # noqa: BLE001 - any failure selects the torch.distributed exchange below on ALL ranks
            print(f"[bench] rank {rank}: native comm unavailable ({exc}); using torch.distributed", file=sys.stderr)
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=f"cuda:{local_rank}")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        native_comm = bool(flag.item())
    if world > 1:
        prov_stream = torch.cuda.ExternalStream(p.stream(), device=f"cuda:{local_rank}")
        comm_stream = torch.cuda.current_stream()
        sum_slots = torch.zeros(2, dtype=torch.float64, device=f"cuda:{local_rank}")
        slot_ptr = [sum_slots[i:i + 1].data_ptr() for i in range(2)]
        slot_view = [sum_slots[i:i + 1] for i in range(2)]
        copied = [torch.cuda.Event(), torch.cuda.Event()]
        reduced = [torch.cuda.Event(), torch.cuda.Event()]
        for e in reduced:
            e.record(comm_stream)
    step_no = [0]

    def step():
        hC = p.fused_elementwise(ew_shader, [hA, hB, hOne], shape, ELEMS)
        hS = p.fused_reduction(red_shader, [hA, hB], (1, 1), ELEMS, 1)
        if native_comm:
            hG = p.comm_allreduce_sum(hS)  # global sum of this step; overlaps the next step's kernels
            p.free(hS)
            return hC, hG
        if world > 1:
            slot = step_no[0] & 1
            step_no[0] += 1
            prov_stream.wait_event(reduced[slot])
            p.copy_to_device(hS, slot_ptr[slot], 1)
            copied[slot].record(prov_stream)
            comm_stream.wait_event(copied[slot])
            dist.all_reduce(slot_view[slot])
            reduced[slot].record(comm_stream)
        return hC, hS

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
    p.timer_begin()
    last_sum = None
    in_flight = []  # N>1: a step's global sum is consumed (freed) SUM_DEPTH steps later, so its all-reduce has that much slack
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
    if native_comm:
        p.comm_fence()  # the timed region ends when the last all-reduce has landed
    elif world > 1:
        prov_stream.wait_stream(comm_stream)  # the timed region ends when the last all-reduce has landed
    ms = p.timer_end_ms()
    sync_all()
    clocks = sampler.stop() if rank == 0 else None
    launches = p.telemetry_snapshot().kernel_launches - t_before
    checksum = float(p.download(last_sum)[0, 0])
    checksum_local = checksum
    p.free(last_sum)
    if world > 1 and not native_comm:
        checksum = float(sum_slots[(step_no[0] - 1) & 1].item())  # the all-reduced sum of the last step
    if world > 1:
        tt = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    value = BYTES_STEP * world * args.steps / (ms * 1e-3) / 1e9

    # ---- roofline of the dominant kernel (fused elementwise, 24 B/elem): isolated launches, CUDA events ------------
    def time_kernel(fn, reps):
        for _ in range(3):
            p.free(fn())
        p.synchronize()
        p.timer_begin()
        for _ in range(reps):
            p.free(fn())  # stream-ordered free: the pool hands the same block back, as in the step loop
        return p.timer_end_ms() / reps

    reps = max(args.steps, 10)
    ew_ms = time_kernel(lambda: p.fused_elementwise(ew_shader, [hA, hB, hOne], shape, ELEMS), reps)
    red_ms = time_kernel(lambda: p.fused_reduction(red_shader, [hA, hB], (1, 1), ELEMS, 1), reps)
    peak, peak_src = measured_peaks()
    achieved = BYTES_EW / (ew_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "rm_fused_ew (C = sin(A).*B+1, 24 B/elem)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src, "us_per_launch": ew_ms * 1e3,
                "reduction_kernel": {"kernel": "rm_fused_red (sum(sin(A).*B+1), 16 B/elem)", "achieved": BYTES_RED / (red_ms * 1e-3) / 1e9,
                                     "frac": BYTES_RED / (red_ms * 1e-3) / 1e9 / peak, "us_per_launch": red_ms * 1e3}}
    profs = sorted((ROOT / "profiles").glob("r*_traffic.json"))
    prof = profs[-1] if profs else ROOT / "profiles" / "none"
    if prof.exists():
        try:
            roofline["traffic"] = json.loads(prof.read_text()).get("rm_fused_ew_dram_bytes_per_launch")
        except Exception:
            pass

    # ---- e2e: the same step through the C ABI with HOST buffers (pinned), H2D + D2H inside the timed region ----------
    hostA, hostB, hostC = pinned_empty(ELEMS), pinned_empty(ELEMS), pinned_empty(ELEMS)
    hostA[:] = A
    hostB[:] = B

    E2E_CHUNKS = 8
    chunk = ELEMS // E2E_CHUNKS
    cshape = (chunk, 1)

    def e2e_step():
        """Same step from HOST buffers through the C ABI, chunked and software-pipelined: the upload of chunk i+1 (H2D stream)
        is enqueued before the download of chunk i (compute stream), so H2D, kernels and D2H overlap on the full-duplex link."""
        pending = None
        partials = []
        for i in range(E2E_CHUNKS + 1):
            nxt = None
            if i < E2E_CHUNKS:
                off = i * chunk * 8
                a = p.upload_ptr(hostA.ctypes.data + off, cshape)
                b = p.upload_ptr(hostB.ctypes.data + off, cshape)
                nxt = (i, a, b)
            if pending is not None:
                j, pa, pb, pc = pending
                p.download_async_into_ptr(pc, hostC.ctypes.data + j * chunk * 8, chunk)
                for h in (pa, pb, pc):
                    p.free(h)
            if nxt is not None:
                i_, a, b = nxt
                c = p.fused_elementwise(ew_shader, [a, b, hOne], cshape, chunk)
                partials.append(p.fused_reduction(red_shader, [a, b], (1, 1), chunk, 1))
                pending = (i_, a, b, c)
            else:
                pending = None
        p.synchronize()                      # C is now complete in host memory
        val = 0.0
        for h in partials:
            val += p.read_scalar(h, 0)
            p.free(h)
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
           "d2h_bytes_per_step": ELEMS * 8 + 8, "ms_per_step": e2e_ms / e2e_steps, "ms_per_step_mean": e2e_mean, "ms_per_step_all": [round(x, 3) for x in per_step],
           "steps": e2e_steps, "statistic": "median over per-step host wall times (each step ends in a stream synchronize)",
           "note": "A,B uploaded from pinned host memory, fused elementwise + fused sum, C and the sum downloaded; 8 chunks, software-pipelined (H2D stream / compute+D2H stream)",
           "checksum_rel_diff_vs_resident": (abs(e2e_val - checksum_local) / abs(checksum_local)) if world == 1 else None}

    # ---- other configs of BASELINE.json, reported beside the headline (not the metric) --------------------------------
    extra = {}
    if not args.no_extra:
        try:
            extra = extra_workloads(p, rank, world, local_rank, dist, torch, CudaArray)
        except Exception as e:  # an optional workload must never take the headline line down
            extra = {"error": str(e)}

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
            "config": {"workload": WORKLOAD,
                       "bytes_per_step_per_gpu": BYTES_STEP, "l2": "inputs (2x128 MiB) + output (128 MiB) exceed the 126 MB L2; no flush needed",
                       "parallelism": (f"{world} independent batches, one NCCL all-reduce of 1 f64 per step "
                                       f"({'provider communicator, rm_comm_allreduce_sum' if native_comm else 'torch.distributed'})") if world > 1 else "single GPU",
                       "host_numa_binding": numa},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "checksum": checksum, "extra": extra,
        }
        print(json.dumps(line))
    p.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def extra_workloads(p, rank, world, local_rank, dist, torch, CudaArray):
    """configs[2] (matmul 8192^3, single GPU) and configs[4] (Monte-Carlo 1e8 x 256, sharded + one all-reduce)."""
    out = {}
    # Monte-Carlo: paths sharded in contiguous ranges; RNG addressing is global, result independent of `world`
    M, T = 100_000_000, 256
    lo, hi = rank * M // world, (rank + 1) * M // world
    drift, scale = (0.05 - 0.5 * 0.2 ** 2) / 252.0, 0.2 * math.sqrt(1.0 / 252.0)
    hS0 = p.fill((hi - lo, 1), 100.0)
    p.set_rng_state(0)
    for it in range(2):
        p.set_rng_state(0)
        p.synchronize()
        if world > 1:
            dist.barrier()
        p.timer_begin()
        hS = p.stochastic_evolution_sharded(hS0, drift, scale, T, lo, M)
        hP = p.payoff_partial_sum(hS, 100.0)
        if world > 1:
            p.synchronize()
            ptr, n = p.device_ptr(hP)
            dist.all_reduce(torch.as_tensor(CudaArray(ptr, n), device=f"cuda:{local_rank}"))
            torch.cuda.synchronize()
        ms = p.timer_end_ms()
        price = p.read_scalar(hP, 0) / M * math.exp(-0.05 * T / 252.0)
        p.free(hS)
        p.free(hP)
    if world > 1:
        tt = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    p.free(hS0)
    out["monte_carlo"] = {"paths": M, "steps": T, "ms": ms, "path_steps_per_s": M * T / (ms * 1e-3), "price": price, "scaling": "strong",
                          "collective": "one NCCL all-reduce of 1 f64" if world > 1 else None}
    if world == 1:
        # configs[3] per-GPU share: [8,2160,3840] f32 normalise pipeline (image_normalize) + 5x5 imfilter on one 4K RGB frame
        from runmat_b200 import B200Provider, ImageNormalizeDescriptor

        with B200Provider(local_rank, device_id=rank + 1000, precision="f32") as p32:
            Bi, H, W = 8, 2160, 3840
            img = np.random.default_rng(3).random(Bi * H * W, dtype=np.float32)
            hI = p32.upload(img, (Bi, H, W))
            d = ImageNormalizeDescriptor(Bi, H, W, 1e-6, gain=1.0123, bias=-0.02, gamma=1.8)
            for _ in range(2):
                p32.free(p32.image_normalize(hI, d))
            p32.synchronize()
            p32.timer_begin()
            for _ in range(5):
                p32.free(p32.image_normalize(hI, d))
            ms = p32.timer_end_ms() / 5
            px = Bi * H * W
            out["image_normalize_8x4k_f32"] = {"ms": ms, "gb_per_s_at_12B_per_px": 12 * px / (ms * 1e-3) / 1e9,
                                               "gb_per_s_at_8B_per_px_algorithmic": 8 * px / (ms * 1e-3) / 1e9}
            for _ in range(2):
                p32.free(p32.reduce_mean_nd(hI, [1, 2]))
            p32.synchronize()
            p32.timer_begin()
            for _ in range(5):
                p32.free(p32.reduce_mean_nd(hI, [1, 2]))
            ms = p32.timer_end_ms() / 5
            out["mean_dims23_8x4k_f32"] = {"ms": ms, "gb_per_s": 4 * px / (ms * 1e-3) / 1e9, "kernel": "rm_fused_red (Strided layout, 8 slices)"}
            p32.free(hI)
            frame = np.random.default_rng(4).random(H * W * 3, dtype=np.float32)
            g = np.exp(-((np.arange(5) - 2)[:, None] ** 2 + (np.arange(5) - 2)[None, :] ** 2) / 2.0)
            hF, hK = p32.upload(frame, (H, W, 3)), p32.upload((g / g.sum()).astype(np.float32))
            for _ in range(2):
                p32.free(p32.imfilter(hF, hK, padding="replicate"))
            p32.synchronize()
            p32.timer_begin()
            for _ in range(5):
                p32.free(p32.imfilter(hF, hK, padding="replicate"))
            ms = p32.timer_end_ms() / 5
            out["imfilter_5x5_4k_rgb_f32"] = {"ms": ms, "gb_per_s_at_8B_per_sample": 8 * H * W * 3 / (ms * 1e-3) / 1e9}
        n = 8192
        rng = np.random.default_rng(7)
        hA = p.upload(rng.uniform(-1, 1, n * n), (n, n))
        hB = p.upload(rng.uniform(-1, 1, n * n), (n, n))
        hC = p.matmul(hA, hB)
        p.free(hC)
        p.synchronize()
        p.timer_begin()
        reps = 2
        for _ in range(reps):
            p.free(p.matmul(hA, hB))
        ms = p.timer_end_ms() / reps
        out["matmul_8192_f64"] = {"ms": ms, "gflops": 2.0 * n ** 3 / (ms * 1e-3) / 1e9,
                                  "engine": "auto -> tcgen05 (Ozaki int8 split, 7 slices = 28 exact int8 GEMMs, TMEM int32 accumulate, f64 recombine)"}
        p.set_matmul_engine(1)
        p.free(p.matmul(hA, hB))
        p.synchronize()
        p.timer_begin()
        p.free(p.matmul(hA, hB))
        ms1 = p.timer_end_ms()
        p.set_matmul_engine(0)
        out["matmul_8192_f64_dmma"] = {"ms": ms1, "gflops": 2.0 * n ** 3 / (ms1 * 1e-3) / 1e9, "engine": "FP64 DMMA mma.sync.m8n8k4"}
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
        out["mldivide_4096_x64"] = {"ms": ms, "gflops_lu": (2.0 / 3.0) * ns ** 3 / (ms * 1e-3) / 1e9}
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


if __name__ == "__main__":
    sys.exit(main())
