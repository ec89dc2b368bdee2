"""Host-side sharding logic for the independent-sample workloads.

One process per GPU (torchrun). The path has NO data-path collective: units (Monte-Carlo paths, images of a batch)
are split into contiguous ranges, every rank evolves/normalises its own range, and the only exchange is the final
sum — ONE all-reduce of a handful of f64 (NCCL over NVLink on GPUs; gloo in the CPU tests of this logic).
"""
from __future__ import annotations

from typing import Tuple


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) of `total` units for `rank`; ranges tile [0, total) exactly and differ by at most 1 unit."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    return rank * total // world, (rank + 1) * total // world


def shard_even_pairs(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Like shard_range but cut points are even, so a Box-Muller pair (elements 2j, 2j+1) is never split across
    ranks (not required for correctness — the sharded kernel handles split pairs — but it avoids computing a pair twice)."""
    lo, hi = shard_range(total, rank, world)
    lo -= lo % 2
    if rank != world - 1:
        hi -= hi % 2
    return lo, hi


def allreduce_sum(tensor, dist=None):
    """Sum-all-reduce in place. `dist` is torch.distributed (already initialised) or None for a single process."""
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM)
    return tensor


def batch_slices_for_rank(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Images [b0, b1) of a [B,H,W] batch owned by `rank`. Batch is the stride-1 axis in RunMat's layout, so the host
    performs the strided split when it uploads a rank's [b1-b0, H, W] tensor."""
    return shard_range(batch, rank, world)


def bind_process_to_gpu_numa(pci_bus_id: str):
    """One process per GPU: run this process (and first-touch its pinned host buffers) on the NUMA node the GPU hangs off, so
    H2D/D2H copies do not cross the inter-socket link. Returns {"node": n, "cpus": k} or None when the topology is not exposed
    (containers without /sys, single-node hosts). Never raises: placement is an optimisation, not a requirement."""
    import os

    try:
        bus = pci_bus_id.lower()
        if len(bus.split(":")[0]) == 8:  # cudaDeviceGetPCIBusId may print an 8-digit domain; sysfs uses 4
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            if "-" in part:
                lo, hi = part.split("-")
                cpus.update(range(int(lo), int(hi) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return {"node": node, "cpus": len(cpus)}
    except Exception:
        return None
