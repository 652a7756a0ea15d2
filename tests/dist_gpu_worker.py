"""Worker of tests/test_gpu_dist.py: run under `python -m torch.distributed.run --nproc-per-node N` (one rank per GPU,
NCCL). SURVEY section 4 item 5: the same batch traced at N GPUs must give identical per-ray monitor rows and identical
merged monitor data as on one GPU. Also drives the C-ABI merge (optb_comm_init / optb_monitor_merge)."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _sorted_rows(res, n_rows):
    from optable_b200 import _abi as A

    cols = {k: res[k][:n_rows].cpu().numpy() for k in A.HIT_I32 + A.HIT_U32 + A.HIT_F64 if k in res}
    key = np.lexsort((cols["hit_pop"].view(np.uint32), cols["hit_monitor"], cols["hit_root"].view(np.uint32)))
    return {k: v[key] for k, v in cols.items()}


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    import optable_b200 as ob
    from optable_b200 import _abi as A
    from optable_b200 import backend
    from optable_b200.workloads import WORKLOADS

    for name, n, limit in (("c2_4f_telescope", 300_001, None), ("c3_doublets_16wl", 100_003, None),
                           ("c5_ripa_64", 20_001, {"max_trace_num": 40})):
        w = WORKLOADS[name]
        sc = w.scene(ob)
        table = ob.OpticalTable()
        table.add_components(sc.components)
        table.add_monitors(sc.monitors)
        bundle = w.bundle(n, 0)
        kw = dict(hit_capacity=n * 70) if name.startswith("c5") else {}
        res = table.trace_bundle(bundle, limit, group=True, gather=True, record_hist=True, **kw)
        lo, hi = res["shard"]
        assert (hi - lo) in (n // world, n // world + 1)
        if rank == 0:
            one = table.trace_bundle(bundle, limit, record_hist=True, **kw)   # the whole bundle on one GPU
            n_one = int(one["counters"][A.C_HITS])
            assert int(res["counters"][A.C_HITS]) == n_one == len(res["hit_root"]), (res["counters"], n_one)
            a, b = _sorted_rows(res, n_one), _sorted_rows(one, n_one)
            for k in a:
                assert np.array_equal(a[k], b[k]), f"{name}: column {k} differs between {world} GPUs and 1 GPU"
            assert torch.equal(res["hist_y"], one["hist_y"]) and torch.equal(res["hist_yz"], one["hist_yz"])
            for c in (A.C_SEGMENTS, A.C_INTERACTIONS, A.C_HITS, A.C_TESTS, A.C_DROPPED, A.C_STATUS):
                assert int(res["counters"][c]) == int(one["counters"][c]), (name, c, res["counters"], one["counters"])
            assert int(res["counters_local"][A.C_INTERACTIONS]) < int(res["counters"][A.C_INTERACTIONS])
            print(f"DIST_CASE_OK {name}: {n} rays on {world} GPUs, {n_one} monitor rows identical to the 1-GPU trace", flush=True)
        dist.barrier()

    # ---- the C-ABI merge a C host would use (include/optb.h: optb_comm_* / optb_monitor_merge) ----
    L = backend.lib()
    L.optb_comm_unique_id.argtypes = [C.c_void_p, C.c_void_p]
    L.optb_comm_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    L.optb_monitor_merge.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.optb_comm_destroy.argtypes = [C.c_void_p]
    eng = backend.Engine.get(local)
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = (C.c_char * 128)()
        eng._check(L.optb_comm_unique_id(eng._ctx, buf))
        uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
    uid = uid.cuda()
    dist.broadcast(uid, 0)
    raw = bytes(uid.cpu().numpy().tobytes())
    eng._check(L.optb_comm_init(eng._ctx, raw, rank, world))
    nm = 3
    hy = torch.full((nm, A.HIST_BINS), rank + 1, dtype=torch.int64, device="cuda")
    hyz = torch.full((nm, A.HIST_BINS, A.HIST_BINS), 10 * (rank + 1), dtype=torch.int64, device="cuda")
    cnt = torch.zeros(A.C_COUNT, dtype=torch.int64, device="cuda")
    cnt[A.C_INTERACTIONS], cnt[A.C_STATUS], cnt[A.C_GENERATIONS] = 100 + rank, (1 if rank == 0 else 4), 5 + rank
    st = torch.cuda.current_stream().cuda_stream
    eng._check(L.optb_monitor_merge(eng._ctx, hy.data_ptr(), hyz.data_ptr(), nm, cnt.data_ptr(), C.c_void_p(st)))
    torch.cuda.synchronize()
    tot = world * (world + 1) // 2
    assert int(hy.min()) == int(hy.max()) == tot and int(hyz.max()) == 10 * tot
    assert int(cnt[A.C_INTERACTIONS]) == 100 * world + world * (world - 1) // 2
    assert int(cnt[A.C_STATUS]) == (5 if world > 1 else 1) and int(cnt[A.C_GENERATIONS]) == 5 + world - 1
    eng._check(L.optb_comm_destroy(eng._ctx))
    if rank == 0:
        print("DIST_CABI_OK optb_monitor_merge", flush=True)
        print("DIST_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
