"""CPU tests of the multi-GPU host logic (wgsparkl_b200/sharded.py): slab partitioning and the neighbour exchange
protocol over torch.distributed with the gloo backend, world_size 2 and 3."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from wgsparkl_b200 import scenes
from wgsparkl_b200.sharded import (INT_MAX, INT_MIN, exchange_with_neighbours, neighbours, particle_block_x,
                                   partition_slabs)


def test_block_x_matches_reference_rounding():
    """ties-to-even: x/h = 0.5 -> round 0 -> cell -1 -> block -1; 1.5 -> round 2 -> cell 1 -> block 0 (grid.wgsl:284-292)."""
    x = np.float32([0.5, 1.5, 2.5, -0.5, 4.6, -3.6])
    assert particle_block_x(x, 1.0, 3).tolist() == [-1, 0, 0, -1, 1, -2]


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_partition_is_contiguous_and_balanced(world):
    s = scenes.sand_dam_3d(160, 6, 6, jitter=True)
    bx = particle_block_x(s["particles"]["position"][:, 0], s["cell_width"], 3)
    slabs = partition_slabs(bx, world)
    assert len(slabs) == world and slabs[0][0] == INT_MIN and slabs[-1][1] == INT_MAX
    counts = []
    for r, (lo, hi) in enumerate(slabs):
        assert lo < hi
        if r:
            assert lo == slabs[r - 1][1]
        counts.append(int(((bx >= lo) & (bx < hi)).sum()))
    assert sum(counts) == len(bx)
    if world <= 4:
        assert max(counts) <= 1.5 * len(bx) / world + 8 * 100  # balanced up to one block column


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 64
        send_l = torch.full((n,), 10 * rank + 1, dtype=torch.uint8)
        send_r = torch.full((n,), 10 * rank + 2, dtype=torch.uint8)
        recv_l = torch.zeros(n, dtype=torch.uint8)
        recv_r = torch.zeros(n, dtype=torch.uint8)
        for _ in range(3):  # repeated substeps reuse the same buffers
            exchange_with_neighbours(dist, send_l, send_r, recv_l, recv_r, rank, world)
        left, right = neighbours(rank, world)
        ok = True
        if left is not None:
            ok &= bool((recv_l == 10 * left + 2).all())  # what the -x neighbour sent towards +x
        else:
            ok &= bool((recv_l == 0).all())
        if right is not None:
            ok &= bool((recv_r == 10 * right + 1).all())
        else:
            ok &= bool((recv_r == 0).all())
        imp = torch.arange(96, dtype=torch.int32) * (rank + 1)
        dist.all_reduce(imp, op=dist.ReduceOp.SUM)
        ok &= bool((imp == torch.arange(96, dtype=torch.int32) * (world * (world + 1) // 2)).all())
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_neighbour_exchange_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000 + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(r, True) for r in range(world)]
