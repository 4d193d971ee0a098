"""CPU suite, part 3: the N>1 host logic on gloo, world_size 2 (no GPU).

The pixel path has no inter-GPU exchange: frames are sharded
frame j -> rank j mod N, and the only collective is one broadcast of the filter taps.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench

    # (1) the single broadcast: rank 0 owns the taps
    taps = bench.broadcast_taps(np.array([16, 64, 96, 64, 16], np.int32) if rank == 0 else None, device="cpu")
    # (2) frame sharding
    mine = bench.shard_frames(10, rank, world)
    # (3) max-over-ranks timing and total count
    t = bench.reduce_max(float(rank + 1), device="cpu")
    total = bench.reduce_sum(float(len(mine)), device="cpu")
    out.put((rank, taps.tolist(), mine, t, total))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_broadcast():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, taps0, mine0, t0, tot0), (r1, taps1, mine1, t1, tot1) = res
    assert taps0 == taps1 == [16, 64, 96, 64, 16]
    assert mine0 == [0, 2, 4, 6, 8] and mine1 == [1, 3, 5, 7, 9]
    assert t0 == t1 == 2.0 and tot0 == tot1 == 10.0


def test_shard_frames_covers_everything_once():
    import bench

    for n in (1, 7, 64, 256):
        for world in (1, 2, 4, 8):
            got = sorted(sum((bench.shard_frames(n, r, world) for r in range(world)), []))
            assert got == list(range(n))
            # BASELINE configs 4 / 5: 256 / 64 frames over the ranks
            assert sum(bench.frames_for_rank(n, r, world) for r in range(world)) == n
    assert [bench.frames_for_rank(256, r, 8) for r in range(8)] == [32] * 8
    assert [bench.frames_for_rank(64, r, 4) for r in range(4)] == [16] * 4
    # both arms of bench.py describe the workload with the same config object
    assert bench.config_dict(2, 32) == bench.config_dict(2, 32) and "workload" in bench.config_dict(1, 32)
