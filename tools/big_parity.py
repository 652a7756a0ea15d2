"""One-off large GPU-vs-oracle parity sweep (run on the GPU box: python tools/big_parity.py). Results: profiles/r1_parity_sweep.md."""
import sys, time, numpy as np
sys.path.insert(0, ".")
import bench
from optable_b200.backend import Engine
from oracle import oracle as O
from oracle import ref_harness as RH
from tests import parity
e = Engine.get(0)
for wl, n in (("c2_4f_telescope", 3_000_000), ("c3_doublets_16wl", 1_500_000), ("c5_ripa_64", 60_000)):
    flat = bench.build_scene(wl)
    arrs = bench.make_bundle(n, 5_000_000, wl).materialise()
    mt = bench._workloads()[wl][2]["max_trace_num"]
    scene = e.upload(flat)
    t0 = time.time(); got = e.trace_arrays(scene, arrs, max_trace_num=mt, max_live=16 * n); t1 = time.time()
    want = O.trace(flat, arrs, max_trace_num=mt, nthreads=16); t2 = time.time()
    a, b = RH.arrays_from_result(want), RH.arrays_from_result(got)
    ok = len(a["seg_root"]) == len(b["seg_root"]) and len(a["hit_root"]) == len(b["hit_root"])
    print(wl, n, "segments", len(a["seg_root"]), len(b["seg_root"]), "hits", len(a["hit_root"]), len(b["hit_root"]), f"gpu {t1-t0:.1f}s oracle {t2-t1:.1f}s")
    if ok:
        bad_leaf = int((a["seg_leaf"] != b["seg_leaf"]).sum()); bad_pop = int((a["seg_pop"] != b["seg_pop"]).sum())
        print("  leaf mismatches", bad_leaf, "pop mismatches", bad_pop)
        try:
            errs = parity.compare(a, b, q_rtol=1e-6, label=wl); print("  max errs", {k: float(f"{v:.2e}") for k, v in errs.items()})
        except AssertionError as ex:
            print("  COMPARE FAIL", str(ex)[:300])
    else:
        # per-root segment counts to locate the differing roots
        ca = np.bincount(a["seg_root"], minlength=n); cb = np.bincount(b["seg_root"], minlength=n)
        bad = np.nonzero(ca != cb)[0]
        print("  roots with different segment counts:", len(bad), bad[:10], ca[bad[:10]], cb[bad[:10]])
