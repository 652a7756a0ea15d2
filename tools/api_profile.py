#!/usr/bin/env python
"""cProfile of one small object-level ray_tracing call (where does a GUI slider move's latency go)."""
import cProfile, os, pstats, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import optable_b200 as ob
from tests import scenes

name = sys.argv[1] if len(sys.argv) > 1 else "telescope_4f"
sc = scenes.REGISTRY[name](ob)
table = ob.OpticalTable()
table.add_components(sc.components)
table.add_monitors(sc.monitors)
for _ in range(3):
    table.ray_tracing(sc.rays, perfomance_limit=sc.limit)
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    table.rays = []
    table.ray_tracing(sc.rays, perfomance_limit=sc.limit)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(28)
