#!/usr/bin/env python
"""Latency of the object-level API (what a GUI slider move pays): build table + ray_tracing on small scenes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import optable_b200 as ob
from tests import scenes

for name in ("gaussian_beam", "doublet", "telescope_4f", "misc_components", "ripa"):
    best = 1e9
    for rep in range(4):
        sc = scenes.REGISTRY[name](ob)
        table = ob.OpticalTable()
        table.add_components(sc.components)
        table.add_monitors(sc.monitors)
        t0 = time.perf_counter()
        out = table.ray_tracing(sc.rays, perfomance_limit=sc.limit)
        best = min(best, time.perf_counter() - t0)
    print(f"{name:18s} {len(sc.rays):3d} rays -> {len(out):5d} segments: ray_tracing {best*1e3:8.2f} ms")

# throughput of the object-level API: many Ray objects through one small scene
import numpy as np
sc = scenes.REGISTRY["telescope_4f"](ob)
r0 = sc.rays[0]
rng = np.random.default_rng(0)
for n in (1000, 20000):
    rays = []
    for k in range(n):
        r = r0.copy()
        r.origin = np.array([r0.origin[0], rng.uniform(-0.3, 0.3), rng.uniform(-0.3, 0.3)])
        rays.append(r)
    best = 1e9
    for rep in range(3):
        table = ob.OpticalTable()
        table.add_components(sc.components)
        table.add_monitors(sc.monitors)
        t0 = time.perf_counter()
        out = table.ray_tracing(rays, perfomance_limit=sc.limit)
        best = min(best, time.perf_counter() - t0)
    print(f"telescope_4f {n:6d} Ray objects -> {len(out):6d} segments: ray_tracing {best*1e3:8.2f} ms "
          f"({len(out)/best:.3g} segments/s incl. building the Ray objects)")
