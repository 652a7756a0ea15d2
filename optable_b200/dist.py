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
