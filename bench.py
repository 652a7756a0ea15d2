#!/usr/bin/env python
"""Benchmark of the optable ray_tracing hot path on B200 (contract: see the repo brief / DESIGN.md section 6).

  python bench.py --gpus N --steps K --warmup W          # CUDA arm (one rank per GPU under torchrun for N > 1)
  python bench.py --impl reference --steps K --warmup W  # CPU arm: oracle port of the reference, all host threads

Workload (BASELINE.json configs[1]): two LENS-9 parametric aspheres as a 4f relay, monitors at
x = 0 and x = 2F1 + 2F2, 1e7 synthetic collimated Gaussian rays per GPU (disc radius 3, splitmix64 positions).
A step = one trace of the whole batch incl. monitor row capture. Metric = ray-surface interactions per second
(an interaction = one popped ray that hit a surface = one output segment of finite length).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ray-surface interactions/sec (fp64)"
UNIT = "interactions/s"
WORKLOAD = "c2_4f_telescope"
BYTES_PER_INTERACTION = 208  # read + write one 104-B ray record per interaction (wavefront form)
# fp64 flops per interaction of this workload, from the ncu instruction counts of profiles/ 
E2E_EXTRA = {}
FLOPS_PER_INTERACTION = 1390  # executed (2*DFMA + DMUL + DADD) / interactions, profiles/r1_trace_kernel_summary.md (r1f)


def _workloads():
    """name -> (scene builder, bundle maker(n, start), params). c2 is the benchmark of record (BASELINE configs[1]);
    the others size up the remaining configs of SURVEY 8(d) for information (`--workload`)."""
    import numpy as np

    import optable_b200 as ob
    from optable_b200.bundle import RayBundle, uniform01
    from tests import scenes

    def c2_scene():
        return scenes.telescope_4f(ob, n_rays=0)

    def c2_rays(n, start):
        return RayBundle.collimated_disc(n, start=start, x0=-10.0, radius=3.0, wavelength=780e-7, w0=61e-4)

    def c3_scene():
        mk = lambda x: ob.Doublet([x, 0, 0], CT1=1.359, CT2=0.6, R1=18.405, R2=-13.734, R3=-39.933,
                                  n12=ob.Glass_NBK7(), n23=ob.Glass_NSF5(), diameter=7.5)
        return scenes.Scene([mk(30.3964), mk(90.3964)], [], [ob.Monitor([150, 0, 0], 10, 10)])

    def c3_rays(n, start):
        b = RayBundle.collimated_disc(n, start=start, x0=0.0, radius=3.0, wavelength=780e-7, w0=61e-4)
        wl = np.linspace(400e-7, 1100e-7, 16)[(np.arange(start, start + n) % 16)]
        b.columns["wavelength"] = wl
        b.columns["q_im"] = np.pi * 61e-4 ** 2 / wl
        return b

    def c4_scene():
        return scenes.cavity(ob, 0.0, 0.0)

    def c4_rays(n, start):
        idx = np.arange(start, start + n, dtype=np.uint64)
        u = [uniform01(idx, k) for k in range(4)]
        d = np.stack([np.ones(n), (2 * u[2] - 1) * 2e-5, (2 * u[3] - 1) * 2e-5], 1)
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        b = RayBundle.collimated_disc(n, start=start, x0=2.0, wavelength=780e-7, w0=61e-4)
        b.columns.update(oy=2 * u[0] - 1, oz=2 * u[1] - 1, dx=d[:, 0].copy(), dy=d[:, 1].copy(), dz=d[:, 2].copy())
        return b

    def c5_scene():
        return scenes.ripa(ob, n_rays=0)

    def c5_rays(n, start):
        p = scenes.ripa(ob, n_rays=0).params
        idx = np.arange(start, start + n, dtype=np.uint64)
        u = [uniform01(idx, k) for k in range(4)]
        o, d0, w0 = p["origin"], p["direction"], p["R1w0"]
        d = np.stack([np.full(n, d0[0]), d0[1] + (2 * u[2] - 1) * 1e-3, d0[2] + (2 * u[3] - 1) * 1e-3], 1)
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        b = RayBundle.collimated_disc(n, start=start, x0=o[0], wavelength=p["wavelength"], w0=w0)
        b.columns.update(oy=o[1] + (2 * u[0] - 1) * w0, oz=o[2] + (2 * u[1] - 1) * w0,
                         dx=d[:, 0].copy(), dy=d[:, 1].copy(), dz=d[:, 2].copy())
        return b

    return {
        "c2_4f_telescope": (c2_scene, c2_rays, dict(max_trace_num=2000, rays=10_000_000, rows_per_ray=2)),
        "c3_doublets_16wl": (c3_scene, c3_rays, dict(max_trace_num=2000, rays=10_000_000, rows_per_ray=1)),
        "c4_cavity_4000": (c4_scene, c4_rays, dict(max_trace_num=4001, rays=1_000_000, rows_per_ray=0)),
        "c5_ripa_64": (c5_scene, c5_rays, dict(max_trace_num=64, rays=1_000_000, rows_per_ray=64, max_live=8)),
    }


def build_scene(workload=WORKLOAD):
    from optable_b200.flatten import FlatScene

    sc = _workloads()[workload][0]()
    return FlatScene(sc.components, sc.monitors)


def make_bundle(n, start, workload=WORKLOAD):
    return _workloads()[workload][1](n, start)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()

    def _nvml_loop(self):
        """NVML in-process: a sample every 10 ms (the nvidia-smi binary needs ~0.3 s per query, i.e. one sample
        for a 0.3 s timed region)."""
        import pynvml as N

        N.nvmlInit()
        h = N.nvmlDeviceGetHandleByIndex(self.index)
        reasons_fn = getattr(N, "nvmlDeviceGetCurrentClocksEventReasons", None) or N.nvmlDeviceGetCurrentClocksThrottleReasons
        bits = (0x8, 0x40, 0x20, 0x4)  # HwSlowdown, HwThermalSlowdown, SwThermalSlowdown, SwPowerCap
        mx = N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM)
        while not self._halt.is_set():
            sm, mask = N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM), int(reasons_fn(h))
            self.rows.append([str(sm), str(mx)] + ["Active" if mask & b else "Not Active" for b in bits])
            self._halt.wait(0.01)

    def run(self):
        try:
            self._nvml_loop()
            return
        except Exception:
            pass
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._halt.wait(0.05)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in self.rows for k in range(4) if len(r) >= 6 and r[2 + k].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def cpu_arm(flat, n_sample, threads, repeats=1, workload=WORKLOAD):
    """Oracle port of the reference on the host cores: interactions/s on the first n_sample rays of the workload."""
    from oracle import oracle as O

    arrs = make_bundle(n_sample, 0, workload).materialise()
    max_trace = _workloads()[workload][2]["max_trace_num"]
    best = None
    inter = 0
    for _ in range(repeats):
        t0 = time.perf_counter()
        # same outputs as the CUDA step: every monitor row + histograms (capacity given: no counting pass)
        out = O.trace(flat, arrs, max_trace_num=max_trace, record_segments=False, record_hits=True, record_hist=True,
                      nthreads=threads, hit_capacity=n_sample * _workloads()[workload][2]["rows_per_ray"] + 1024)
        dt = time.perf_counter() - t0
        inter = int(out["counters"][1])
        best = dt if best is None else min(best, dt)
    return inter / best, inter, best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    flat = build_scene(args.workload)
    threads = os.cpu_count() or 1
    n_sample = args.ref_rays
    for _ in range(args.warmup):
        cpu_arm(flat, max(n_sample // 10, 100), threads, workload=args.workload)
    total, dt = 0, 0.0
    for _ in range(args.steps):  # only the trace is timed (cpu_arm builds the synthetic rays outside its timer)
        _, inter, step_s = cpu_arm(flat, n_sample, threads, workload=args.workload)
        total += inter
        dt += step_s
    value = total / dt
    sample = f"first {n_sample} rays of the {args.workload} batch per step, oracle/optb_oracle.c (C port of the reference), {threads} threads"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "rays_per_step": n_sample, "surfaces": int(flat.n_leaves), "monitors": int(flat.n_monitors)},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def bind_to_gpu_numa_node(index):
    """Pin this rank (and the pinned host buffers it allocates next) to the CPUs closest to its GPU, so that the
    host<->device copies of eight ranks do not all cross one socket link. Best effort."""
    try:
        import pynvml

        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(index))
    except Exception:
        pass


def run_cuda(args):
    import torch
    import torch.distributed as dist

    from optable_b200 import _abi as A
    from optable_b200.backend import Engine
    from optable_b200.bundle import DeviceTrace
    from optable_b200.flatten import rays_struct

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    dev = f"cuda:{local}"
    engine = Engine.get(local)
    flat = build_scene(args.workload)
    wprm = _workloads()[args.workload][2]
    n = args.rays or wprm["rays"]
    bundle = make_bundle(n, rank * n, args.workload)  # weak scaling: every rank traces its own n rays of the endless bundle
    rays_dev = bundle.to_torch(device=dev)
    hit_cap = n * wprm["rows_per_ray"] + 1024
    dt = DeviceTrace(engine, flat, n, hit_cap, record_hist=True, max_trace_num=wprm["max_trace_num"], chain_len=args.chain_len)
    stream = torch.cuda.current_stream()

    def merge_monitors():
        # the only cross-GPU exchange of the path: monitor histograms (+ row counts); rows stay sharded
        if world > 1:
            dist.all_reduce(dt.t["hist_y"])
            dist.all_reduce(dt.t["hist_yz"])

    live = wprm.get("max_live", 0) * n or None  # live-ray budget of the wavefront (splitting scenes)

    def step():
        dt.run(rays_dev, live)
        merge_monitors()

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    cnt = dt.counters()
    engine._raise_status(cnt)
    inter_per_step, hits_per_step = int(cnt[A.C_INTERACTIONS]), int(cnt[A.C_HITS])
    launches_per_step = int(cnt[A.C_LAUNCHES])
    # ---- timed region: device-resident inputs, CUDA events on the launching stream, max over ranks ----
    sampler = ClockSampler(local)
    sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    kern_ms = []
    ev[0].record(stream)
    for k in range(args.steps):
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record(stream)
        dt.run(rays_dev, live)
        k1.record(stream)
        merge_monitors()
        ev[k + 1].record(stream)
        kern_ms.append((k0, k1))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    total_ms = ev[0].elapsed_time(ev[-1])
    trace_ms = float(np.mean([a.elapsed_time(b) for a, b in kern_ms]))
    clocks = sampler.stop()
    if world > 1:
        tm = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        total_ms = float(tm.item())
        ti = torch.tensor([inter_per_step], device=dev, dtype=torch.int64)
        dist.all_reduce(ti)
        inter_all = int(ti.item())
    else:
        inter_all = inter_per_step
    ms_per_step = total_ms / args.steps
    value = inter_all / (ms_per_step * 1e-3)

    # ---- end to end through the C ABI with HOST buffers (optb_trace_host): H2D rays + D2H monitor rows ----
    hit_columns, hit_dtypes, row_bytes = dt.hit_columns, {k: dt.t[k].dtype for k in dt.hit_columns}, dt.hit_row_bytes()
    hist_shapes = (tuple(dt.t["hist_y"].shape), tuple(dt.t["hist_yz"].shape))
    for k in list(dt.t):  # the device-resident result buffers are not needed any more: give the memory back
        if k.startswith(("hit_", "seg_")):
            del dt.t[k]
    del rays_dev
    engine._workspace = None
    torch.cuda.empty_cache()
    if args.e2e_steps <= 0:
        e2e_value, h2d, d2h, e2e_steps = None, 0, 0, 0
    else:
        e2e_value, h2d, d2h, e2e_steps = run_e2e(args, engine, dt, bundle, n, hit_cap, hit_columns, hit_dtypes, row_bytes,
                                                 hist_shapes, inter_per_step, hits_per_step, inter_all, world, dev)
    finish(args, engine, flat, world, rank, n, inter_per_step, hits_per_step, launches_per_step, ms_per_step, trace_ms,
           value, clocks, e2e_value, h2d, d2h, e2e_steps)
    if world > 1:
        dist.destroy_process_group()


def run_e2e(args, engine, dt, bundle, n, hit_cap, hit_columns, hit_dtypes, row_bytes, hist_shapes, inter_per_step,
            hits_per_step, inter_all, world, dev):
    import torch
    import torch.distributed as dist

    from optable_b200 import _abi as A
    from optable_b200.flatten import rays_struct

    host = bundle.to_torch(pin=True)
    host_np = {k: v.numpy() for k, v in host.items()}
    host_np["length"] = None
    rs = rays_struct({**{k: None for k in A.RAY_F64}, **host_np})
    rs.n = n
    res = A.Result()
    res.seg_capacity, res.hit_capacity = 0, hit_cap
    host_out = {}
    for k in hit_columns:
        host_out[k] = torch.empty(hit_cap, dtype=hit_dtypes[k]).pin_memory()
        setattr(res, k, host_out[k].data_ptr())
    hy = torch.zeros(hist_shapes[0], dtype=torch.int64).pin_memory()
    hyz = torch.zeros(hist_shapes[1], dtype=torch.int64).pin_memory()
    hc = torch.zeros(A.C_COUNT, dtype=torch.int64).pin_memory()
    res.hist_y, res.hist_yz, res.counters = hy.data_ptr(), hyz.data_ptr(), hc.data_ptr()
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    wprm = _workloads()[args.workload][2]
    for _ in range(2):
        engine.trace_host(dt.scene, rs, dt.prm, res)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        engine.trace_host(dt.scene, rs, dt.prm, res)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        te = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_s = float(te.item())
    assert int(hc[A.C_STATUS]) == 0 and int(hc[A.C_INTERACTIONS]) == inter_per_step, (hc.tolist(), inter_per_step)
    e2e_value = inter_all * e2e_steps / e2e_s
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    d2h = hits_per_step * row_bytes + hy.numel() * 8 + hyz.numel() * 8 + A.C_COUNT * 8
    # for information: the same call when the caller only wants the monitors' histograms back (no row columns)
    import copy

    prm2 = copy.copy(dt.prm)
    prm2.record_hits = 0
    res2 = A.Result()
    res2.hist_y, res2.hist_yz, res2.counters = hy.data_ptr(), hyz.data_ptr(), hc.data_ptr()
    engine.trace_host(dt.scene, rs, prm2, res2)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        engine.trace_host(dt.scene, rs, prm2, res2)
    torch.cuda.synchronize()
    hist_s = time.perf_counter() - t0
    if world > 1:
        te = torch.tensor([hist_s], device=dev, dtype=torch.float64)
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        hist_s = float(te.item())
    E2E_EXTRA["e2e_histograms_only"] = {"value": inter_all * e2e_steps / hist_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                                        "d2h_bytes_per_step": int(hy.numel() * 8 + hyz.numel() * 8 + A.C_COUNT * 8)}
    return e2e_value, h2d, d2h, e2e_steps


def finish(args, engine, flat, world, rank, n, inter_per_step, hits_per_step, launches_per_step, ms_per_step, trace_ms,
           value, clocks, e2e_value, h2d, d2h, e2e_steps):
    if rank != 0:
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak, peak_src = (peaks.get("hbm_gbs"), "measured") if peaks.get("hbm_gbs") else (6650.0, "fallback")
    achieved = BYTES_PER_INTERACTION * inter_per_step / (trace_ms * 1e-3) / 1e9
    # measured DRAM bytes per launch of this workload (ncu dram__bytes_read+write, profiles/r1_trace_kernel_summary.md)
    traffic = 1.6686e9 if (args.workload == WORKLOAD and n == 10_000_000) else None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": traffic, "peak_source": peak_src, "kernel": "trace_kernel", "kernel_ms": trace_ms,
                "note": "kernel is FP64-pipe bound, not HBM bound: see roofline_fp64 and DESIGN.md"}
    fp64_peak = engine.fp64_peak_tflops()
    roofline_fp64 = {"bound": "fp64", "peak": fp64_peak, "unit": "TFLOP/s", "peak_source": "measured in-run (DFMA chain kernel)"}
    if FLOPS_PER_INTERACTION and args.workload == WORKLOAD:
        af = FLOPS_PER_INTERACTION * inter_per_step / (trace_ms * 1e-3) / 1e12
        roofline_fp64.update({"achieved": af, "frac": af / fp64_peak})
    threads = os.cpu_count() or 1
    try:
        os.sched_setaffinity(0, range(threads))  # the CPU baseline gets every host core again
    except Exception:
        pass
    cpu_val, cpu_inter, cpu_dt = cpu_arm(flat, args.cpu_rays, threads, workload=args.workload)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": args.workload, "rays_per_gpu": n, "surfaces": int(flat.n_leaves), "monitors": int(flat.n_monitors),
                       "interactions_per_step_per_gpu": inter_per_step, "monitor_rows_per_step_per_gpu": hits_per_step,
                       "l2": "inputs+outputs per step (%.2f GB) exceed the 126 MB L2" % ((h2d + d2h) / 1e9),
                       "parallelism": f"rays sharded over {world} GPU(s), scene tables replicated"},
            "clocks": clocks, "gpu_launches": launches_per_step * args.steps,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": e2e_steps, "api": "optb_trace_host (C ABI, pinned host buffers)"},
            **E2E_EXTRA, "roofline": roofline, "roofline_fp64": roofline_fp64,
            "cpu_baseline": {"value": cpu_val, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"first {args.cpu_rays} rays of the same batch ({cpu_inter} interactions in {cpu_dt:.2f} s), "
                                       f"oracle/optb_oracle.c with {threads} threads"}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--rays", type=int, default=0, help="rays per GPU per step (0 = the workload's default)")
    ap.add_argument("--workload", default=WORKLOAD, choices=["c2_4f_telescope", "c3_doublets_16wl", "c4_cavity_4000", "c5_ripa_64"])
    ap.add_argument("--cpu-rays", type=int, default=400_000, help="bounded sample for the in-run CPU baseline")
    ap.add_argument("--ref-rays", type=int, default=400_000, help="rays per step of the reference arm")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--chain-len", type=int, default=0, help="max in-register pops per launch (0 = unlimited); scheduling only")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "cuda":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
