#!/usr/bin/env python
"""Host<->device copy bandwidth per rank, alone and with all ranks copying at once (what bounds `e2e` at N > 1).
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_probe.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if os.environ.get("PROBE_BIND", "1") == "1":
    bench.bind_to_gpu_numa_node(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
nbytes = 1 << 30
h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
h.zero_()


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return nbytes * reps / (time.perf_counter() - t0) / 1e9


res = {}
for name, fn in (("d2h", lambda: h.copy_(d, non_blocking=True)), ("h2d", lambda: d.copy_(h, non_blocking=True))):
    # alone: ranks take turns
    for r in range(world):
        if world > 1:
            dist.barrier()
        if r == rank:
            res[name + "_alone"] = timed.__wrapped__(fn) if hasattr(timed, "__wrapped__") else None
    res[name + "_together"] = timed(fn)
# alone measurement without barriers inside
for name, fn in (("d2h", lambda: h.copy_(d, non_blocking=True)), ("h2d", lambda: d.copy_(h, non_blocking=True))):
    for r in range(world):
        if world > 1:
            dist.barrier()
        if r == rank:
            fn(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(5):
                fn()
            torch.cuda.synchronize()
            res[name + "_alone"] = nbytes * 5 / (time.perf_counter() - t0) / 1e9
    if world > 1:
        dist.barrier()
aff = sorted(os.sched_getaffinity(0))
print(f"rank {rank}: " + " ".join(f"{k}={v:.1f}GB/s" for k, v in sorted(res.items()) if v) + f" cpus {aff[0]}..{aff[-1]} ({len(aff)})", flush=True)
if world > 1:
    dist.destroy_process_group()
