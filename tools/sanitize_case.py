#!/usr/bin/env python
"""Small traces that touch every kernel (chain, wavefront with splitting and pop cap, family-serial, large-table
global path) for compute-sanitizer runs: `compute-sanitizer --tool memcheck python tools/sanitize_case.py`."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from optable_b200.backend import Engine  # noqa: E402
from tests import golden_io  # noqa: E402

engine = Engine.get(0)
for name in ("telescope_4f", "misc_components", "caps_binding", "mma_small", "ripa"):
    flat, rays, params, ref = golden_io.load(name)
    scene = engine.upload(flat)
    out = engine.trace_arrays(scene, rays, record_hist=True, **params)
    assert len(out["seg_root"]) == len(ref["seg_root"]), name
    print(name, "ok", len(out["seg_root"]), "segments")
