#!/usr/bin/env python
"""Trace time of an INCOHERENT small scene (tests/scenes.misc_components: beam splitter, block with a hole, cylinder
mirror, wedge, lenses, MLA, DMD; rays spread over the whole aperture, splitting, pop cap 40) -- the counterpart of the
coherent benchmark bundles when a walk change is A/B-tested (OPTB_LIB_PATH selects the library)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import optable_b200 as ob
from optable_b200 import _abi as A
from optable_b200.backend import Engine
from optable_b200.bundle import DeviceTrace, RayBundle
from optable_b200.flatten import FlatScene
from tests import scenes

engine = Engine.get(0)
for name, n, limit in (("misc_components", 400_000, 40), ("mma_small", 400_000, 60)):
    sc = scenes.REGISTRY[name](ob)
    flat = FlatScene(sc.components, sc.monitors)
    if name == "misc_components":
        arrs = scenes.ray_arrays(n, [-3, 0, 0], [0, 15, 0.4], [1, 0, 0], [0, 0.03, 0.03], wavelengths=(633e-7, 500e-7))
    else:
        arrs = scenes.ray_arrays(n, [0, 0, 0], [0, 0.1, 0.03], [1, 0, 0], [0, 0.02, 0.02])
    rays = RayBundle({k: arrs[k] for k in A.RAY_F64}, n).to_torch(device="cuda:0")
    dt = DeviceTrace(engine, flat, n, 40 * n, record_hist=True, max_trace_num=limit)
    live = 40 * n
    for _ in range(2):
        dt.run(rays, live)
    torch.cuda.synchronize()
    ms = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); dt.run(rays, live); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    c = dt.counters()
    print(f"{name}: {np.median(ms):.3f} ms, {int(c[A.C_INTERACTIONS])} interactions, status {int(c[A.C_STATUS])}", flush=True)
    dt.close()
