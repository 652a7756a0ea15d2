#!/usr/bin/env python
"""How ill-conditioned is a random scene's longest path? Traces tests/scenes.fuzz(SEED) twice with the CPU oracle, the
second time with every initial ray moved by ~1e-13, and prints how far the segment origins move per pop. Used to
classify whole-path differences of tools/fuzz_sweep.py (CUDA vs oracle) above the 1e-7 bar: a path that turns 1e-13 into
1e-5 cannot be compared more tightly than that, whatever computes it (profiles/r2_parity_sweep.md).

  python tools/path_sensitivity.py 100889 [--extended]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import optable_b200 as ob
from optable_b200.flatten import pack_rays, trace_cap
from oracle import oracle as O, ref_harness as RH
from tests import scenes

seed = int(sys.argv[1])
sc = scenes.fuzz(ob, seed, n_rays=64, extended="--extended" in sys.argv)
flat = sc.flat()
arrs, fam, unit = pack_rays(sc.rays)
prm = dict(max_trace_num=trace_cap(sc.limit), unit=unit, n_families=len(fam))
a = RH.arrays_from_result(O.trace(flat, arrs, **prm))
moved = {k: (v.copy() if hasattr(v, "copy") else v) for k, v in arrs.items()}
moved["oy"] = moved["oy"] * (1 + 1e-13) + 1e-13
b = RH.arrays_from_result(O.trace(flat, moved, **prm))
print("components:", [type(c).__name__ for c in sc.components])
if len(a["seg_root"]) != len(b["seg_root"]):
    print(f"the perturbed trace has {len(b['seg_root'])} segments instead of {len(a['seg_root'])}: a decision flipped")
    sys.exit(0)
d = np.abs(np.asarray(a["seg_o"]) - np.asarray(b["seg_o"])).max(1)
i = int(d.argmax())
root = a["seg_root"][i]
rows = np.asarray(a["seg_root"]) == root
print(f"largest move of a segment origin: {d.max():.3e} (root {root}, pop {a['seg_pop'][i]}, {rows.sum()} pops) for an input move of 1e-13")
print("per pop of that root:", " ".join(f"{v:.1e}" for v in d[rows]))
