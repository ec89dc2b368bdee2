"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: shard ranges + the single final all-reduce."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_shard_ranges_tile_exactly():
    from runmat_b200.sharding import shard_even_pairs, shard_range

    for total in (0, 1, 7, 8, 100_000_000, 1_000_003):
        for world in (1, 2, 3, 4, 8):
            cuts = [shard_range(total, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1
            ev = [shard_even_pairs(total, r, world) for r in range(world)]
            assert ev[0][0] == 0 and ev[-1][1] == total and all(a[1] == b[0] for a, b in zip(ev, ev[1:]))
            assert all(lo % 2 == 0 for lo, _ in ev)
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _worker(rank, world, port, M, T, q):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist

    from oracle_binding import Oracle
    from runmat_b200.sharding import allreduce_sum, shard_range

    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = Oracle()
    lo, hi = shard_range(M, rank, world)
    partial = orc.mc_lcg_payoff_sum(M, T, lo, hi - lo)  # this rank's sum(max(S-K,0)) over its contiguous path range
    t = torch.tensor([partial], dtype=torch.float64)
    allreduce_sum(t, dist)                               # the path's ONLY collective: 1 f64
    q.put((rank, float(t.item()), lo, hi))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_monte_carlo_sum_matches_single_process(orc):
    import torch.multiprocessing as mp

    M, T, world = 4001, 16, 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, M, T, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    whole = orc.mc_lcg_payoff_sum(M, T, 0, M)
    for _, total, _, _ in results:
        assert abs(total - whole) <= 1e-12 * abs(whole)  # only the order of the final two-term sum differs
    ranges = sorted((lo, hi) for _, _, lo, hi in results)
    assert ranges[0][0] == 0 and ranges[-1][1] == M and ranges[0][1] == ranges[1][0]


def test_lcg_image_shard_matches_oracle_and_strided_split(orc):
    """The per-rank generator equals the oracle's restatement of runmat_lcg.m:59-79, and generating a rank's images directly
    equals the strided host split of the whole batch-fastest tensor."""
    from runmat_b200.sharding import batch_slices_for_rank, lcg_image_shard, split_batch_strided

    B, H, W = 6, 9, 7
    whole = lcg_image_shard(B, H, W, 0, B, seed=3.0)
    assert whole.dtype == np.float32 and whole.flags.f_contiguous
    assert np.array_equal(whole, orc.image_lcg_fill(B, H, W, seed=3.0))
    flat = whole.reshape(-1, order="F")
    assert flat[2 + B * (4 + H * 5)] == whole[2, 4, 5]                       # batch is the stride-1 axis
    for world in (1, 2, 3, 4):
        parts = []
        for r in range(world):
            b0, b1 = batch_slices_for_rank(B, r, world)
            shard = lcg_image_shard(B, H, W, b0, b1 - b0, seed=3.0)
            assert np.array_equal(shard, split_batch_strided(whole, r, world))
            assert np.array_equal(shard, orc.image_lcg_fill(B, H, W, seed=3.0, b0=b0, bcount=b1 - b0))
            assert shard.flags.f_contiguous and shard.shape == (b1 - b0, H, W)
            parts.append(shard)
        assert np.array_equal(np.concatenate(parts, axis=0), whole)
    with pytest.raises(ValueError):
        lcg_image_shard(1 << 20, 4096, 4096, 0, 1)


def _image_worker(rank, world, port, B, H, W, q):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist

    from oracle_binding import Oracle
    from runmat_b200.sharding import allreduce_sum, batch_slices_for_rank, lcg_image_shard

    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = Oracle()
    b0, b1 = batch_slices_for_rank(B, rank, world)
    imgs = lcg_image_shard(B, H, W, b0, b1 - b0)
    out = orc.image_normalize(imgs, 1e-6, gain=1.0123, bias=-0.02, gamma=1.8, f32=True)   # per-image statistics: no exchange
    err = out.astype(np.float32) - imgs
    t = torch.tensor([float(np.sum((err * err).astype(np.float64)))], dtype=torch.float64)
    allreduce_sum(t, dist)                                                               # the path's ONLY collective: 1 f64
    q.put((rank, float(t.item()) / (B * H * W), b0, b1))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_image_batch_mse_matches_single_process(orc):
    """configs[3] host logic at world_size 2 (gloo): batch split by images, per-rank normalise (the oracle stands in for the
    device), squared-error partial, ONE all-reduce -> the MSE of the unsharded batch (runmat_lcg.m:81-94)."""
    import torch.multiprocessing as mp

    from runmat_b200.sharding import lcg_image_shard

    B, H, W, world = 6, 40, 24, 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_image_worker, args=(r, world, port, B, H, W, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    imgs = lcg_image_shard(B, H, W, 0, B)
    out = orc.image_normalize(imgs, 1e-6, gain=1.0123, bias=-0.02, gamma=1.8, f32=True)
    err = out.astype(np.float32) - imgs
    whole = float(np.sum((err * err).astype(np.float64))) / (B * H * W)
    for _, mse, _, _ in results:
        assert abs(mse - whole) <= 1e-12 * abs(whole)
    ranges = sorted((b0, b1) for _, _, b0, b1 in results)
    assert ranges == [(0, 3), (3, 6)]


def test_numa_binding_is_best_effort():
    """Host placement helper never raises and reports None when the device is not in sysfs (as in this container)."""
    from runmat_b200.sharding import bind_process_to_gpu_numa

    assert bind_process_to_gpu_numa("0000:ff:1f.7") is None
    assert bind_process_to_gpu_numa("garbage") is None


def test_reference_arm_under_torchrun_prints_one_json_line():
    """bench.py --impl reference launched the way the driver launches it at N > 1 (torchrun, one process per GPU): rank 0 alone runs
    the CPU reference path and prints exactly ONE JSON line on stdout, the other ranks exit 0 without work; library banners and
    torchrun's own notices must not reach stdout."""
    import json
    import pathlib
    import subprocess
    import sys

    root = pathlib.Path(__file__).resolve().parent.parent
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29541", str(root / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[:2000]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["unit"] == "GB/s" and d["higher_is_better"] is True and d["value"] > 0
