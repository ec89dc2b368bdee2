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


def split_batch_strided(imgs, rank: int, world: int):
    """The strided host split of a `[B,H,W]` batch-fastest tensor (element (b,h,w) at b + B*(h + H*w),
    simple_provider.rs:7937-7938): rank r's images [b0,b1) become their own dense column-major `[b1-b0,H,W]` tensor.
    `imgs` is a numpy array of shape (B,H,W) in Fortran order (or anything reshapeable to it); returns an F-ordered copy."""
    import numpy as np

    B = imgs.shape[0]
    b0, b1 = batch_slices_for_rank(B, rank, world)
    return np.asfortranarray(imgs[b0:b1, ...])


def lcg_image_shard(B: int, H: int, W: int, b0: int, bcount: int, seed: float = 0.0):
    """Images [b0, b0+bcount) of the deterministic batch of benchmarks/4k-image-processing/runmat_lcg.m:59-79, generated directly
    in the rank's own `[bcount,H,W]` layout (no global tensor is ever materialised on a rank):
        idx = (b-1)*H*W + seed + y*W + x;  state = mod(1664525*idx + 1013904223, 2^32);  pixel = single(state)/single(2^32).
    All products stay below 2^53 for B*H*W < 5.4e9, so integer arithmetic reproduces MATLAB's doubles exactly.
    Returns float32, Fortran order."""
    import numpy as np

    if (B * H * W + int(seed)) * 1664525 + 1013904223 >= 2 ** 53:
        raise ValueError("index range exceeds exact double arithmetic; the script's mod() would round")
    out = np.empty((bcount, H, W), dtype=np.float32, order="F")
    yx = (np.arange(H, dtype=np.uint64)[:, None] * np.uint64(W) + np.arange(W, dtype=np.uint64)[None, :])  # [H,W]: y*W + x
    for k in range(bcount):
        idx = yx + np.uint64((b0 + k) * H * W + int(seed))
        state = (np.uint64(1664525) * idx + np.uint64(1013904223)) & np.uint64(0xFFFFFFFF)
        # single(state) ./ single(2^32): round the integer to f32 first, then divide in f32 (exact: power of two)
        out[k, :, :] = state.astype(np.float32) / np.float32(4294967296.0)
    return out


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
