#!/usr/bin/env python
"""Latency of one "slider event" (optable/interact.py:455-457): move one component, trace again.
Compares, on the 7,689-leaf ripa scene and on the 4f telescope, (a) what round 1 did -- flatten the whole scene, upload,
trace -- with (b) FlatScene.refresh + optb_scene_update_nodes + trace and, for scenes that cannot split, (c) the same with
the trace replayed from a CUDA graph. Few rays, as in the GUI."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import optable_b200 as ob
from optable_b200.backend import Engine
from optable_b200.bundle import DeviceTrace, RayBundle
from optable_b200.flatten import FlatScene
from optable_b200.workloads import WORKLOADS, ripa, telescope_4f

engine = Engine.get(0)


def med(f, reps=20):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        f()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return 1e3 * float(np.median(ts))


def case(name, sc, moved, bundle, max_trace, live, graph):
    n = bundle.n
    rays = bundle.to_torch(device="cuda:0")
    rows = n * 64 + 64

    def full():
        moved.TX(1e-4)
        flat = FlatScene(sc.components, sc.monitors)
        dt = DeviceTrace(engine, flat, n, rows, record_hist=True, max_trace_num=max_trace)
        dt.run(rays, live)
        dt.counters()
        dt.scene.close()

    flat = FlatScene(sc.components, sc.monitors)
    dt = DeviceTrace(engine, flat, n, rows, record_hist=True, max_trace_num=max_trace)
    dt.run(rays, live)

    def incremental():
        moved.TX(1e-4)
        dt.scene.update_nodes(flat.refresh(moved))
        dt.run(rays, live)
        dt.counters()

    out = {"scene": name, "leaves": flat.n_leaves, "rays": n, "full_reflatten_ms": med(full, 8), "refresh_update_trace_ms": med(incremental)}
    if graph:
        dt.capture(rays, live)

        def replay():
            moved.TX(1e-4)
            dt.scene.update_nodes(flat.refresh(moved))
            dt.replay()
            dt.counters()

        out["refresh_update_graph_replay_ms"] = med(replay)
    print(out, flush=True)


sc = telescope_4f(ob, n_rays=0)
case("4f telescope", sc, sc.components[1], RayBundle.collimated_disc(1000, radius=2.5), 2000, None, True)
sc = ripa(ob, n_rays=0)
case("ripa", sc, sc.components[0].components[1], WORKLOADS["c5_ripa_64"].bundle(1000, 0), 64, 64 * 1000, False)
