#!/usr/bin/env python
"""One random scene of tools/fuzz_sweep.py under the magnifying glass: the restarted single interactions whose segment
length (or q) differs most between the CUDA engine and the oracle, with the geometry of the leaf that was hit.
  python tools/fuzz_case.py SEED [--extended] [--reference-roots]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import optable_b200 as ob
from optable_b200 import _abi as A
from optable_b200.backend import Engine
from optable_b200.flatten import pack_rays, trace_cap
from oracle import oracle as O, ref_harness as RH
from tests import parity, scenes

seed = int(sys.argv[1])
KW = {"reference_roots": True} if "--reference-roots" in sys.argv else {}
sc = scenes.fuzz(ob, seed, n_rays=64, extended="--extended" in sys.argv)
flat = sc.flat()
arrs, fam, unit = pack_rays(sc.rays)
raw = O.trace(flat, arrs, max_trace_num=trace_cap(sc.limit), unit=unit, n_families=len(fam))
batch = parity.restart_batch(raw, np.nonzero(np.isinf(arrs["length"]))[0])
p1 = dict(max_trace_num=3, unit=unit, n_families=len(batch["ox"]))
e = Engine.get(0)
if "--whole-first" in sys.argv:   # what tools/fuzz_sweep.py does before the restarted batch
    whole = RH.arrays_from_result(e.trace_arrays(e.upload(flat), arrs, max_trace_num=trace_cap(sc.limit), unit=unit, n_families=len(fam), **KW))
    ww = RH.arrays_from_result(raw)
    if len(ww["seg_root"]) == len(whole["seg_root"]):
        r = parity._rel(ww["seg_length"], whole["seg_length"], parity.LENGTH_FLOOR)
        k = int(np.argmax(r))
        print(f"whole paths: worst seg_length rel {r[k]:.3e} at row {k} root {ww['seg_root'][k]} pop {ww['seg_pop'][k]} leaf {ww['seg_leaf'][k]} "
              f"oracle {ww['seg_length'][k]!r} cuda {whole['seg_length'][k]!r}")
want = RH.arrays_from_result(O.trace(flat, batch, **p1))
got = RH.arrays_from_result(e.trace_arrays(e.upload(flat), batch, **p1, **KW))
assert len(want["seg_root"]) == len(got["seg_root"])
lw, lg = np.asarray(want["seg_length"]), np.asarray(got["seg_length"])
fin = np.isfinite(lw) & np.isfinite(lg)
rel = np.zeros_like(lw)
rel[fin] = np.abs(lw[fin] - lg[fin]) / np.maximum(np.abs(lw[fin]), parity.LENGTH_FLOOR)
print("restarted batch:", len(batch["ox"]), "rays,", len(lw), "segments; parity._rel max", float(parity._rel(lw, lg, parity.LENGTH_FLOOR).max()))
leaf_nodes = [i for i in range(flat.n_nodes) if flat.node_i[i, A.NI_GEOM] not in (A.G_GROUP, A.G_GRID)]
names = {getattr(A, k): k for k in dir(A) if k.startswith("G_")}
for i in np.argsort(-rel)[:5]:
    leaf = int(want["seg_leaf"][i])
    node = leaf_nodes[leaf] if 0 <= leaf < len(leaf_nodes) else -1
    g = names.get(int(flat.node_i[node, A.NI_GEOM]), "?") if node >= 0 else "-"
    d = np.asarray(want["seg_d"])[i]
    print(f"row {i}: rel {rel[i]:.3e} length oracle {lw[i]!r} cuda {lg[i]!r} diff {lw[i]-lg[i]:.3e} leaf {leaf} {g} "
          f"pop {want['seg_pop'][i]} d {d} p {flat.node_f[node, A.NF_P:A.NF_P + 8] if node >= 0 else ''}")
