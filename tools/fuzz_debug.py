#!/usr/bin/env python
"""Locate the first CUDA-vs-oracle difference of one fuzz scene: python tools/fuzz_debug.py SEED [n_rays]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import optable_b200 as ob
from optable_b200 import _abi as A
from optable_b200.backend import Engine
from optable_b200.flatten import FlatScene, pack_rays, trace_cap
from oracle import oracle as O
from tests import scenes

seed = int(sys.argv[1]); n_rays = int(sys.argv[2]) if len(sys.argv) > 2 else 64
sc = scenes.fuzz(ob, seed, n_rays=n_rays)
flat = sc.flat()
arrs, fam, unit = pack_rays(sc.rays)
prm = dict(max_trace_num=trace_cap(sc.limit), unit=unit, n_families=len(fam))
want = O.trace(flat, arrs, **prm)
e = Engine.get(0)
got = e.trace_arrays(e.upload(flat), arrs, **prm)
print("components:", [type(c).__name__ for c in sc.components])
print("segments", len(want["seg_root"]), len(got["seg_root"]))
n = min(len(want["seg_root"]), len(got["seg_root"]))
bad = np.nonzero((want["seg_leaf"][:n] != got["seg_leaf"][:n]) | (want["seg_root"][:n] != got["seg_root"][:n]) | (want["seg_pop"][:n] != got["seg_pop"][:n]))[0]
print("first differing rows", bad[:5])
if len(bad):
    k = int(bad[0])
    for name, r in (("oracle", want), ("cuda", got)):
        print(name, "root", r["seg_root"][k], "pop", r["seg_pop"][k], "leaf", r["seg_leaf"][k], "len %.17g" % r["seg_length"][k],
              "o", [float("%.17g" % r["seg_o" + c][k]) for c in "xyz"], "d", [float("%.17g" % r["seg_d" + c][k]) for c in "xyz"])
    o = np.array([want["seg_o" + c][k] for c in "xyz"]); d = np.array([want["seg_d" + c][k] for c in "xyz"])
    leaf_nodes = np.nonzero(flat.node_i[:, A.NI_LEAF] >= 0)[0]
    for lf in sorted({int(want["seg_leaf"][k]), int(got["seg_leaf"][k])} - {-1}):
        node = int(leaf_nodes[lf])
        comp = flat.leaves[lf]
        res = O.intersect(flat, node, o, d)
        print(" leaf", lf, "node", node, type(comp).__name__, type(comp.surface).__name__, "geom", flat.node_i[node, A.NI_GEOM],
              "oracle intersect ->", res)
if not len(bad):
    from oracle import ref_harness as RH
    from tests import parity
    a, b = RH.arrays_from_result(want), RH.arrays_from_result(got)
    rel = parity._rel(a["seg_length"], b["seg_length"], 1e-3)
    for k in np.argsort(rel)[::-1][:4]:
        lf = int(want["seg_leaf"][k])
        comp = flat.leaves[lf] if lf >= 0 else None
        print("row", k, "root", want["seg_root"][k], "pop", want["seg_pop"][k], "leaf", lf, type(comp).__name__, type(comp.surface).__name__ if comp else None,
              "len oracle %.17g cuda %.17g rel %.2e" % (want["seg_length"][k], got["seg_length"][k], rel[k]),
              "o", [float(want["seg_o" + c][k]) for c in "xyz"], "d", [float(want["seg_d" + c][k]) for c in "xyz"])
        if comp is not None:
            node = int(np.nonzero(flat.node_i[:, A.NI_LEAF] == lf)[0][0])
            print("   node_f geom params", flat.node_f[node, A.NF_GEOM:A.NF_GEOM + 8] if hasattr(A, "NF_GEOM") else "", "oracle intersect", O.intersect(flat, node, [want["seg_o" + c][k] for c in "xyz"], [want["seg_d" + c][k] for c in "xyz"]))
