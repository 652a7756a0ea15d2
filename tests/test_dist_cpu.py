"""CPU, world_size 2, gloo: the host-side multi-GPU logic (sharding, monitor merge) of optable_b200.dist."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from optable_b200 import dist as D


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = D.shard_bounds(n, rank, world)
        # every rank "traces" its block: one monitor row per ray with an even global index
        roots = torch.arange(lo, hi)
        keep = roots % 2 == 0
        cols = {"hit_root": (roots[keep] - lo).to(torch.int32), "hit_py": roots[keep].double() * 0.5}
        nrows = int(keep.sum())
        hy = torch.zeros((1, 30), dtype=torch.int64)
        hy[0, rank] = nrows
        hyz = torch.zeros((1, 30, 30), dtype=torch.int64)
        hyz[0, rank, rank] = nrows
        D.merge_histograms(hy, hyz)
        counts = D.gather_row_counts(nrows)
        allrows = D.gather_rows(cols, nrows, root_offset=lo)
        expect = torch.arange(0, n, 2)
        assert sum(counts) == len(expect) == int(hy.sum()) == int(hyz.sum())
        assert torch.equal(allrows["hit_root"].long(), expect)          # rank order == global ray order
        assert torch.equal(allrows["hit_py"], expect.double() * 0.5)
        assert int(hy[0, 0]) + int(hy[0, 1]) == len(expect)
    finally:
        dist.destroy_process_group()


def test_shard_and_merge_world2():
    mp.spawn(_worker, args=(2, _free_port(), 1001), nprocs=2, join=True)


def test_shard_bounds_cover_and_keep_families_whole():
    n = 103
    for world in (1, 2, 4, 8):
        blocks = [D.shard_bounds(n, r, world) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
    fam = np.repeat(np.arange(20), 5)  # 20 families of 5 rays (e.g. 5 wavelengths multiplexed per ray id)
    for world in (2, 3, 8):
        blocks = [D.shard_bounds(len(fam), r, world, fam) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == len(fam)
        for (lo, hi), (lo2, _) in zip(blocks, blocks[1:]):
            assert hi == lo2 and (hi == len(fam) or fam[hi] != fam[hi - 1])
