#!/usr/bin/env python
"""Upper bound on what ray coherence can buy on the ripa scene: trace 1e6 rays with the origin/angle jitter scaled
by `s` (0 = all rays identical -> every lane of a warp walks the same nodes)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from optable_b200 import _abi as A
from optable_b200.backend import Engine
from optable_b200.bundle import DeviceTrace

engine = Engine.get(0)
flat = bench.build_scene("c5_ripa_64")
n = 1_000_000
for s in (1.0, 0.1, 0.0):
    b = bench.make_bundle(n, 0, "c5_ripa_64")
    base = bench.make_bundle(1, 0, "c5_ripa_64")
    from tests import scenes
    import optable_b200 as ob
    p = scenes.ripa(ob, n_rays=0).params
    for k, c0 in (("oy", p["origin"][1]), ("oz", p["origin"][2])):
        b.columns[k] = c0 + s * (b.columns[k] - c0)
    d0 = p["direction"]
    d = np.stack([b.columns["dx"], d0[1] + s * (b.columns["dy"] - d0[1]), d0[2] + s * (b.columns["dz"] - d0[2])], 1)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    b.columns.update(dx=d[:, 0].copy(), dy=d[:, 1].copy(), dz=d[:, 2].copy())
    dt = DeviceTrace(engine, flat, n, n * 64 + 1024, record_hist=True, max_trace_num=64)
    rays = b.to_torch(device="cuda:0")
    for _ in range(3):
        dt.run(rays, 8 * n)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        dt.run(rays, 8 * n)
    e1.record(); torch.cuda.synchronize()
    cnt = dt.counters()
    print(f"jitter x{s}: {e0.elapsed_time(e1)/3:.2f} ms/step, interactions {int(cnt[A.C_INTERACTIONS])}, generations {int(cnt[A.C_GENERATIONS])}")
    del dt
    torch.cuda.empty_cache()
