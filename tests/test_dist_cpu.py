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


def _oracle_tracer(table, bundle, perfomance_limit=None, record_hits=True, record_hist=True, engine=None, **kw):
    """Stand-in for bundle.trace_bundle on the CPU (the C oracle), same result contract with CPU tensors."""
    from optable_b200 import _abi as A
    from optable_b200.flatten import FlatScene, trace_cap
    from oracle import oracle as O

    flat = FlatScene(table.components, table.monitors)
    out = O.trace(flat, bundle.materialise(), max_trace_num=trace_cap(perfomance_limit), record_segments=False,
                  record_hits=record_hits, record_hist=record_hist)
    res = {k: torch.from_numpy(np.ascontiguousarray(out[k]).view(np.int32) if out[k].dtype == np.uint32 else np.ascontiguousarray(out[k]))
           for k in A.HIT_I32 + A.HIT_U32 + A.HIT_F64}
    res["hist_y"], res["hist_yz"] = torch.from_numpy(out["hist_y"].copy()), torch.from_numpy(out["hist_yz"].copy())
    res["counters"] = out["counters"].copy()
    return res


def _table():
    import optable_b200 as ob
    from optable_b200.workloads import telescope_4f

    sc = telescope_4f(ob, n_rays=0)
    t = ob.OpticalTable()
    t.add_components(sc.components)
    t.add_monitors(sc.monitors)
    return t


def _sharded_worker(rank, world, port, n, path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from optable_b200.bundle import RayBundle

        res = D.trace_sharded(_table(), RayBundle.collimated_disc(n), gather=True, tracer=_oracle_tracer)
        if rank == 0:
            np.savez(path, **{k: (v.numpy() if hasattr(v, "numpy") else np.asarray(v)) for k, v in res.items() if k != "shard"})
    finally:
        dist.destroy_process_group()


def test_trace_sharded_world2_equals_single_process(tmp_path):
    """SURVEY section 4 item 5 on the CPU (gloo): the same batch traced by 2 ranks gives identical per-ray monitor
    rows (global keys, rank order = ray order), identical merged histograms and summed counters."""
    from optable_b200 import _abi as A
    from optable_b200.bundle import RayBundle

    n = 3001
    path = str(tmp_path / "sharded.npz")
    mp.spawn(_sharded_worker, args=(2, _free_port(), n, path), nprocs=2, join=True)
    got = np.load(path)
    want = D.trace_sharded(_table(), RayBundle.collimated_disc(n), tracer=_oracle_tracer)   # no process group: whole bundle
    for k in A.HIT_I32 + A.HIT_U32 + A.HIT_F64:
        np.testing.assert_array_equal(got[k], want[k].numpy(), err_msg=k)
    np.testing.assert_array_equal(got["hist_y"], want["hist_y"].numpy())
    np.testing.assert_array_equal(got["hist_yz"], want["hist_yz"].numpy())
    for c in (A.C_INTERACTIONS, A.C_HITS, A.C_TESTS, A.C_STATUS):
        assert int(got["counters"][c]) == int(want["counters"][c])
    assert int(got["counters_local"][A.C_HITS]) < int(got["counters"][A.C_HITS])
