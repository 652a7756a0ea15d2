#!/usr/bin/env python
"""CUDA vs oracle on many random scenes (tests/scenes.fuzz):
  python tools/fuzz_sweep.py FIRST_SEED N_SCENES [n_rays] [--extended] [--caps] [--reference-roots]
(on the GPU so far: 300 --extended scenes, 40 --caps scenes; see profiles/r1_parity_sweep.md)"""
import os, sys, time, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import optable_b200 as ob
from optable_b200.backend import Engine
from optable_b200.flatten import FlatScene, pack_rays, trace_cap
from oracle import oracle as O, ref_harness as RH
from tests import parity, scenes

flags = {a for a in sys.argv[1:] if a.startswith("--")}   # --extended: the whole component zoo; --caps: binding interact caps
argv = [a for a in sys.argv[1:] if not a.startswith("--")]
first, count = int(argv[0]), int(argv[1])
n_rays = int(argv[2]) if len(argv) > 2 else 64
EXTENDED, CAPS = "--extended" in flags, "--caps" in flags
KW = {"reference_roots": True} if "--reference-roots" in flags else {}   # curved roots from the device's brentq restatement
e = Engine.get(0)
pops = rays = hits = 0
flagged, failures, worst = [], [], collections.defaultdict(float)
flagged1, worst1, pops1, rays1 = [], collections.defaultdict(float), 0, 0
kinds = collections.Counter()
t0 = time.time()
for seed in range(first, first + count):
    sc = scenes.fuzz(ob, seed, n_rays=n_rays, caps=CAPS, extended=EXTENDED)
    flat = sc.flat()
    arrs, fam, unit = pack_rays(sc.rays)
    prm = dict(max_trace_num=trace_cap(sc.limit), unit=unit, n_families=len(fam))
    raw = O.trace(flat, arrs, **prm)
    want = RH.arrays_from_result(raw)
    got = RH.arrays_from_result(e.trace_arrays(e.upload(flat), arrs, **prm, **KW))
    for c in sc.components:
        kinds[type(c).__name__] += 1
    try:
        errs, ties = parity.compare_flagging_ties(flat, want, got, rtol=1e-6, q_rtol=10 * max(parity.q_rtol_for(flat), 1e-6), label=f"seed {seed}")
        for k, v in errs.items():
            worst[k] = max(worst[k], float(v))
        flagged += [(seed, r) for r in ties]
        pops += len(want["seg_root"]); hits += len(want["hit_root"]); rays += len(sc.rays)
        batch = parity.restart_batch(raw, np.nonzero(np.isinf(arrs["length"]))[0])
        p1 = dict(max_trace_num=3, unit=unit, n_families=len(batch["ox"]))
        want1 = RH.arrays_from_result(O.trace(flat, batch, **p1))
        got1 = RH.arrays_from_result(e.trace_arrays(e.upload(flat), batch, **p1, **KW))
        errs1, ties1 = parity.compare_flagging_ties(flat, want1, got1, q_rtol=1e-5 if parity.q_rtol_for(flat) > parity.RTOL else parity.RTOL, label=f"seed {seed} restarted")
        for k, v in errs1.items():
            worst1[k] = max(worst1[k], float(v))
        flagged1 += [(seed, r) for r in ties1]
        pops1 += len(want1["seg_root"]); rays1 += len(batch["ox"])
    except AssertionError as ex:
        failures.append((seed, str(ex)[:200]))
        continue
print(f"{count} scenes, {rays} rays, {pops} pops, {hits} monitor rows in {time.time()-t0:.0f} s")
print("flagged (tie) roots:", len(flagged), flagged[:20])
print("failures:", len(failures), failures[:5])
import re
by_kind = collections.defaultdict(list)   # value-tolerance failures by field; anything else (a count / index mismatch) as "DECISION"
for seed, msg in failures:
    m = re.match(r"seed \d+( restarted)?: (\w+) max relative error ([0-9.e+-]+)", msg)
    by_kind[(m.group(2) + (m.group(1) or "")) if m else "DECISION"].append((float(m.group(3)) if m else 0.0, seed))
print("failures by kind (count, worst value, its seed):", {k: (len(v), *max(v)) for k, v in by_kind.items()})
print("worst relative errors:", {k: float(f"{v:.2e}") for k, v in worst.items()})
print(f"restarted single interactions: {rays1} rays, {pops1} pops, flagged ties {len(flagged1)}")
print("worst relative errors (restarted, bar 1e-9; q 1e-5 in scenes with FD-curvature aspheres):", {k: float(f"{v:.2e}") for k, v in worst1.items()})
print("components drawn:", dict(kinds))
