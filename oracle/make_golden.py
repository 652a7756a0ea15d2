"""Generate tests/golden/*.npz from the REAL reference (run in the build container only).

    python -m oracle.make_golden

Each fixture holds, for one scene of tests/scenes.py built with the reference's own classes:
the flattened tables (optable_b200.flatten.FlatScene.to_arrays), the packed input rays, the trace
parameters, and the reference's results (oracle.ref_harness.run_reference): every output segment in
(root, pop) order with the winning leaf index, and every Monitor._data_raw row.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_harness as RH  # noqa: E402
from optable_b200.flatten import FlatScene, pack_rays, trace_cap  # noqa: E402
from tests import scenes  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def make(name):
    ref = RH.load_reference()
    sc = scenes.REGISTRY[name](ref)
    flat = FlatScene(sc.components, sc.monitors)
    arrs, fam_ids, unit = pack_rays(sc.rays)
    r = RH.run_reference(sc)
    leaves = r.pop("_leaves")
    # interact counts after the trace, per cap slot and family
    caps = np.zeros((max(flat.n_capslots, 1), len(fam_ids)), np.int32)
    for s, comp in enumerate(flat.capslots):
        for f, rid in enumerate(fam_ids):
            caps[s, f] = comp._interact_count.get(rid, 0)
    d = {}
    d.update({"scene_" + k: v for k, v in flat.to_arrays().items()})
    d.update({"ray_" + k: v for k, v in arrs.items()})
    d["param_max_trace_num"] = np.int64(trace_cap(sc.limit))
    d["param_unit"] = np.float64(unit)
    d["param_n_families"] = np.int64(len(fam_ids))
    d.update({"ref_" + k: v for k, v in r.items()})
    d["ref_cap_counts"] = caps
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    return len(r["seg_root"]), len(r["hit_root"])


if __name__ == "__main__":
    for name in (sys.argv[1:] or scenes.REGISTRY):
        nseg, nhit = make(name)
        print(f"{name:20s} segments={nseg} hits={nhit}")
