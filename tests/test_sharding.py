"""CPU tests of the multi-GPU partition logic, including a world_size-2 gloo run (the path bench.py takes
under torchrun: every rank works on its own shard, only timings are reduced)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from smrf_b200 import sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_split_range_covers_exactly():
    for n in (0, 1, 7, 100, 10000):
        for parts in (1, 2, 3, 8):
            pieces = [sharding.split_range(n, parts, i) for i in range(parts)]
            assert pieces[0][0] == 0 and pieces[-1][1] == n
            for (a, b), (c, d) in zip(pieces[:-1], pieces[1:]):
                assert b == c and a <= b
            sizes = [b - a for a, b in pieces]
            assert max(sizes) - min(sizes) <= 1


def test_time_blocks_and_row_slabs():
    blocks = [b for r in range(8) for b in sharding.time_blocks(8760, 720, 8, r)]
    covered = np.zeros(8760, dtype=int)
    for a, b in blocks:
        covered[a:b] += 1
    assert (covered == 1).all()
    slabs = [s for r in range(4) for s in sharding.row_slabs(10000, 4, r, rows_per_slab=1000)]
    covered = np.zeros(10000, dtype=int)
    for a, b in slabs:
        assert b - a <= 1000
        covered[a:b] += 1
    assert (covered == 1).all()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, T, ny):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # each rank "distributes" its own shards: here it just marks them; no data-path collective
        mine_t = np.zeros(T, dtype=np.int64)
        for a, b in sharding.time_blocks(T, 96, world, rank):
            mine_t[a:b] += 1
        mine_r = np.zeros(ny, dtype=np.int64)
        for a, b in sharding.row_slabs(ny, world, rank, rows_per_slab=64):
            mine_r[a:b] += 1
        tt, rr = torch.from_numpy(mine_t), torch.from_numpy(mine_r)
        dist.all_reduce(tt)
        dist.all_reduce(rr)
        assert bool((tt == 1).all()) and bool((rr == 1).all())
        # the bench's timing reduction: max over ranks
        ms = torch.tensor([10.0 + rank], dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        assert ms.item() == 10.0 + world - 1
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_partition():
    port = _free_port()
    mp.spawn(_worker, args=(2, port, 1000, 333), nprocs=2, join=True)
