#!/usr/bin/env python
"""A/B timing of liboptb variants (tools/build_variants.sh) on the bench workloads: one subprocess per variant
(OPTB_LIB_PATH selects the library), device-resident inputs, CUDA events, median of `--steps` steps.

  python tools/variant_bench.py --variants r1,new --workloads c2_4f_telescope,c3_doublets_16wl
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(args):
    import numpy as np
    import torch

    import bench
    from optable_b200 import _abi as A
    from optable_b200.backend import Engine
    from optable_b200.bundle import DeviceTrace

    engine = Engine.get(0)
    out = {}
    for wl in args.workloads.split(","):
        w = bench.workloads()[wl]
        flat = w.flat()
        n = int((args.rays or min(w.rays_per_gpu, 10_000_000 if w.max_live == 0 else 1_000_000)) * args.scale)
        bundle = w.bundle(n, 0)
        rays_dev = bundle.to_torch(device="cuda:0")
        dt = DeviceTrace(engine, flat, n, n * w.rows_per_ray + 1024, record_hist=True, max_trace_num=w.max_trace_num,
                         hit_columns=bench.E2E_COLUMNS)
        live = w.max_live * n or None
        for _ in range(3):
            dt.run(rays_dev, live)
        torch.cuda.synchronize()
        cnt = dt.counters()
        ms = []
        for _ in range(args.steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dt.run(rays_dev, live)
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        out[wl] = {"ms": float(np.median(ms)), "min_ms": float(min(ms)), "interactions": int(cnt[A.C_INTERACTIONS]),
                   "hits": int(cnt[A.C_HITS]), "tests": int(cnt[A.C_TESTS]), "status": int(cnt[A.C_STATUS]),
                   "box_tests": int(cnt[A.C_BOX_TESTS]) if len(cnt) > 9 else None, "launches": int(cnt[A.C_LAUNCHES])}
        del dt, rays_dev
        engine._workspace = None
        torch.cuda.empty_cache()
    print("RESULT " + json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variants", default="")
    ap.add_argument("--workloads", default="c2_4f_telescope,c3_doublets_16wl,c4_cavity_4000,c5_ripa_64")
    ap.add_argument("--steps", type=int, default=7)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--rays", type=int, default=0)
    ap.add_argument("--child", action="store_true")
    args = ap.parse_args()
    if args.child:
        return child(args)
    rows = {}
    for v in args.variants.split(","):
        env = dict(os.environ)
        if v != "intree":
            env["OPTB_LIB_PATH"] = os.path.join(ROOT, "build_variants", f"liboptb_{v}.so")
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", "--workloads", args.workloads, "--steps", str(args.steps),
                            "--scale", str(args.scale), "--rays", str(args.rays)], env=env, capture_output=True, text=True)
        line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
        if not line:
            print(f"{v}: FAILED\n{r.stdout[-2000:]}\n{r.stderr[-3000:]}")
            continue
        rows[v] = json.loads(line[0][7:])
        print(v, " ".join(f"{w.split('_')[0]}={d['ms']:.3f}ms" for w, d in rows[v].items()), flush=True)
    ref = None
    for v, d in rows.items():
        sig = {w: (x["interactions"], x["hits"], x["status"]) for w, x in d.items()}
        if ref is None:
            ref = sig
        elif sig != ref:
            print(f"WARNING: {v} counters differ from the first variant: {sig} vs {ref}")
    print("VARIANTS " + json.dumps(rows))


if __name__ == "__main__":
    main()
