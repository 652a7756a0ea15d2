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
