"""Multi-GPU plumbing for the ray_tracing path: shard the initial rays, merge the monitors.

The path shards without any data-path collective: every rank traces a contiguous block of the
initial rays against its own replica of the scene tables. The only exchange is the monitor merge at the end:
all-reduce of the histograms, all-gather of the row counts and, when one rank wants every row, a gather-v of
the row columns. torch.distributed is the transport (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n: int, rank: int, world: int, family=None):
    """Contiguous block [lo, hi) of initial-ray indices for `rank`. Concatenating the blocks in rank order
    keeps the reference's (initial ray, pop) output order. When `family` (dense Ray._id index per ray) is
    given, boundaries move forward to the next family change so that rays sharing an `_id` (interact-cap
    state, optical_component.py:136-149) stay on one rank; `family` must then be grouped contiguously."""
    lo, hi = n * rank // world, n * (rank + 1) // world
    if family is not None and n:
        fam = np.asarray(family)

        def snap(i):
            while 0 < i < n and fam[i] == fam[i - 1]:
                i += 1
            return i

        lo, hi = snap(lo), snap(hi)
    return lo, hi


def merge_histograms(hist_y, hist_yz, group=None):
    """In-place sum over ranks of the per-monitor histograms (int64 tensors)."""
    import torch.distributed as dist

    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(hist_y, group=group)
        dist.all_reduce(hist_yz, group=group)
    return hist_y, hist_yz


def gather_row_counts(n_rows: int, device="cpu", group=None):
    """Number of monitor rows every rank holds (list of ints, rank order)."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return [int(n_rows)]
    mine = torch.tensor([int(n_rows)], dtype=torch.int64, device=device)
    out = [torch.zeros_like(mine) for _ in range(dist.get_world_size(group))]
    dist.all_gather(out, mine, group=group)
    return [int(t.item()) for t in out]


def gather_rows(columns: dict, n_rows: int, root_offset: int = 0, group=None):
    """Gather-v of monitor row columns: every rank receives the rows of all ranks, concatenated in rank order
    (= initial-ray block order). `hit_root` is shifted by `root_offset` (the rank's first global ray index) so
    that row keys stay global. Columns are 1-D tensors with at least n_rows valid leading entries."""
    import torch
    import torch.distributed as dist

    cols = {k: v[:n_rows] for k, v in columns.items()}
    if "hit_root" in cols and root_offset:
        cols["hit_root"] = cols["hit_root"] + int(root_offset)
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return cols
    any_col = next(iter(cols.values()))
    counts = gather_row_counts(n_rows, any_col.device, group)
    width = max(counts)
    out = {}
    for k, v in cols.items():
        padded = torch.zeros(width, dtype=v.dtype, device=v.device)
        padded[:n_rows] = v
        parts = [torch.zeros_like(padded) for _ in counts]
        dist.all_gather(parts, padded, group=group)
        out[k] = torch.cat([p[:c] for p, c in zip(parts, counts)])
    return out


def merge_counters(counters, group=None):
    """Sum over ranks of the OPTB_C_* counters of one sharded trace (numpy int64 array or tensor); the status word is
    a bit mask and is OR-ed. Returns a numpy array."""
    import torch
    import torch.distributed as dist

    from . import _abi as A

    c = torch.as_tensor(np.asarray(counters, dtype=np.int64)).clone()
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return c.numpy()
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    bits = torch.tensor([(int(c[A.C_STATUS]) >> k) & 1 for k in range(16)], dtype=torch.int64)
    c[A.C_STATUS] = 0
    gens = c[A.C_GENERATIONS].clone()
    both = torch.cat([c, bits, gens.reshape(1)]).to(dev)
    dist.all_reduce(both, group=group)
    mx = gens.reshape(1).to(dev)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    both = both.cpu()
    out = both[:len(c)].clone()
    out[A.C_STATUS] = sum((1 << k) for k in range(16) if int(both[len(c) + k]) > 0)
    out[A.C_GENERATIONS] = int(mx.item())  # generations run side by side on the ranks: the longest shard counts
    return out.numpy()


def trace_sharded(table, bundle, perfomance_limit=None, group=None, gather=False, record_hits=True, record_hist=True,
                  engine=None, tracer=None, **kw):
    """Multi-GPU form of `OpticalTable.trace_bundle`: rank r of `group` traces the contiguous block
    [lo, hi) = shard_bounds(bundle.n, r, world) of the initial rays on its own GPU against its own replica of the
    scene tables; no data-path collective. The only exchange is the monitor merge at the end: all-reduce of the
    histograms and of the counters and, with `gather=True`, a gather-v of the monitor row columns so that every
    rank holds all rows (concatenated in rank order = initial-ray block order). Row keys are global (`hit_root` /
    `hit_key` count from ray 0 of the whole bundle).

    Returns the dict of `trace_bundle` with merged `hist_y` / `hist_yz` / `counters`, plus `counters_local` and
    `shard` = (lo, hi). Without an initialised process group it is `trace_bundle` on the whole bundle.
    `tracer(table, bundle, perfomance_limit, record_hits=, record_hist=, engine=, **kw)` defaults to the CUDA
    `bundle.trace_bundle`; the CPU tests of this module pass a stand-in."""
    import torch.distributed as dist

    from . import _abi as A

    on = dist.is_initialized()
    rank, world = (dist.get_rank(group), dist.get_world_size(group)) if on else (0, 1)
    lo, hi = shard_bounds(bundle.n, rank, world)
    if tracer is None:
        from .bundle import trace_bundle as tracer
    res = tracer(table, bundle.slice(lo, hi), perfomance_limit, record_hits=record_hits, record_hist=record_hist,
                 engine=engine, **kw)
    rows = {k: v for k, v in res.items() if k.startswith("hit_")}
    n_rows = int(res["counters"][A.C_HITS]) if record_hits else 0
    if lo:
        if "hit_root" in rows:
            rows["hit_root"] = rows["hit_root"] + int(lo)     # uint32 bit pattern in int32: wraps like unsigned
        if "hit_key" in rows:
            rows["hit_key"] = rows["hit_key"] + (int(lo) << 32)
    out = dict(res)
    out["counters_local"] = np.asarray(res["counters"]).copy()
    out["shard"] = (lo, hi)
    if on and world > 1:
        if record_hist:
            merge_histograms(res["hist_y"], res["hist_yz"], group)
        out["counters"] = merge_counters(res["counters"], group)
        if gather and record_hits:
            rows = gather_rows(rows, n_rows, 0, group)
    out.update(rows)
    return out
