"""ctypes wrapper around oracle/_ref/liboptb_oracle.so (the C restatement, oracle/optb_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs.
The product package (optable_b200/) must never import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from optable_b200 import _abi as A
from optable_b200.flatten import FlatScene, rays_struct

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "liboptb_oracle.so")
_lib = None


def build(force=False):
    """Compile the checker with the committed Makefile (gcc only)."""
    srcs = [os.path.join(_HERE, "optb_oracle.c"), os.path.join(_HERE, "Makefile"),
            os.path.join(os.path.dirname(_HERE), "include", "optb.h")]   # (the header carries the ABI version both sides check)
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(f) for f in srcs):
        subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.optb_oracle_trace.restype = C.c_int
        L.optb_oracle_trace.argtypes = [C.POINTER(A.SceneDesc), C.POINTER(A.Rays), C.POINTER(A.Params),
                                        C.POINTER(A.Result), C.c_int]
        L.optb_oracle_count.restype = C.c_int
        L.optb_oracle_count.argtypes = [C.POINTER(A.SceneDesc), C.POINTER(A.Rays), C.POINTER(A.Params),
                                        C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.optb_oracle_intersect.restype = C.c_int
        L.optb_oracle_intersect.argtypes = [C.POINTER(A.SceneDesc), C.c_int, C.c_void_p, C.c_void_p, C.c_double,
                                            C.c_void_p, C.c_void_p]
        L.optb_oracle_material_n.restype = C.c_double
        L.optb_oracle_material_n.argtypes = [C.POINTER(A.SceneDesc), C.c_int, C.c_double]
        L.optb_oracle_hist_bin.restype = C.c_int
        L.optb_oracle_hist_bin.argtypes = [C.c_double, C.c_double, C.c_double]
        L.optb_oracle_slab.restype = C.c_int
        L.optb_oracle_slab.argtypes = [C.c_void_p] * 5
        _lib = L
    return _lib


def make_params(max_trace_num=2000, unit=1e-2, record_segments=True, record_hits=True, record_hist=False,
                chain_len=0, n_families=1):
    p = A.Params()
    p.max_trace_num = int(max_trace_num)
    p.unit = float(unit)
    p.record_segments, p.record_hits, p.record_hist = int(record_segments), int(record_hits), int(record_hist)
    p.chain_len, p.n_families = int(chain_len), int(n_families)
    return p


def alloc_result(n_seg, n_hit, n_monitors, n_capslots, n_families, cap_counts=None):
    """Host result arrays + the ctypes struct viewing them."""
    arrs = {}
    for k in A.SEG_F64:
        arrs[k] = np.zeros(n_seg, np.float64)
    for k in A.SEG_U32:
        arrs[k] = np.zeros(n_seg, np.uint32)
    for k in A.SEG_I32:
        arrs[k] = np.zeros(n_seg, np.int32)
    for k in A.HIT_I32:
        arrs[k] = np.zeros(n_hit, np.int32)
    for k in A.HIT_U32:
        arrs[k] = np.zeros(n_hit, np.uint32)
    for k in A.HIT_F64:
        arrs[k] = np.zeros(n_hit, np.float64)
    arrs["hist_y"] = np.zeros((max(n_monitors, 1), A.HIST_BINS), np.int64)
    arrs["hist_yz"] = np.zeros((max(n_monitors, 1), A.HIST_BINS, A.HIST_BINS), np.int64)
    if cap_counts is None:
        cap_counts = np.zeros((max(n_capslots, 1), max(n_families, 1)), np.int32)
    arrs["cap_counts"] = np.ascontiguousarray(cap_counts, dtype=np.int32)
    arrs["counters"] = np.zeros(A.C_COUNT, np.int64)
    r = A.Result()
    r.seg_capacity, r.hit_capacity = n_seg, n_hit
    for k, a in arrs.items():
        setattr(r, k, a.ctypes.data)
    r._keepalive = arrs
    return r, arrs


def trace(flat: FlatScene, ray_arrs, max_trace_num=2000, unit=1e-2, record_segments=True, record_hits=True,
          record_hist=False, n_families=None, cap_counts=None, nthreads=1, seg_capacity=None, hit_capacity=None):
    """Run the CPU restatement. Returns a dict of numpy arrays trimmed to the produced rows, in
    reference order: segments (root, pop); hits (root, monitor, pop)."""
    L = lib()
    n = len(ray_arrs["ox"])
    if n_families is None:
        fam = ray_arrs.get("family")
        n_families = int(fam.max()) + 1 if fam is not None and n else max(n, 1)
    desc = flat.desc()
    rs = rays_struct(ray_arrs)
    prm = make_params(max_trace_num, unit, record_segments, record_hits, record_hist, 0, n_families)
    n_seg = n_hit = 0
    known = (not record_segments or seg_capacity is not None) and (not record_hits or hit_capacity is not None)
    if known:  # capacities given by the caller: no counting pass (used when this run is being timed)
        n_seg, n_hit = int(seg_capacity or 0) if record_segments else 0, int(hit_capacity or 0) if record_hits else 0
    elif record_segments or record_hits:
        scratch = None
        if flat.n_capslots:
            scratch = np.array(cap_counts if cap_counts is not None
                               else np.zeros((flat.n_capslots, n_families), np.int32), dtype=np.int32, copy=True)
        cs, ch = C.c_int64(0), C.c_int64(0)
        rc = L.optb_oracle_count(C.byref(desc), C.byref(rs), C.byref(prm),
                                 None if scratch is None else scratch.ctypes.data, C.byref(cs), C.byref(ch))
        assert rc == 0
        n_seg, n_hit = (cs.value if record_segments else 0), (ch.value if record_hits else 0)
    res, arrs = alloc_result(n_seg, n_hit, flat.n_monitors, flat.n_capslots, n_families, cap_counts)
    rc = L.optb_oracle_trace(C.byref(desc), C.byref(rs), C.byref(prm), C.byref(res), int(nthreads))
    if rc != 0:
        raise RuntimeError(f"optb_oracle_trace failed: {rc}")
    arrs["counters"][A.C_SEGMENTS] = n_seg if record_segments else -1
    if known and record_hits:  # trim to the rows produced
        rows = int(arrs["counters"][A.C_HITS])
        for k in A.HIT_I32 + A.HIT_U32 + A.HIT_F64:
            arrs[k] = arrs[k][:rows]
    return arrs


def intersect(flat: FlatScene, node: int, o, d, length=np.inf):
    L = lib()
    desc = flat.desc()
    o = np.ascontiguousarray(o, np.float64)
    d = np.ascontiguousarray(d, np.float64)
    P = np.zeros(3)
    t = np.zeros(1)
    hit = L.optb_oracle_intersect(C.byref(desc), int(node), o.ctypes.data, d.ctypes.data, float(length),
                                  P.ctypes.data, t.ctypes.data)
    return (P, float(t[0])) if hit else (None, None)


def material_n(flat: FlatScene, m: int, wl_m: float) -> float:
    return float(lib().optb_oracle_material_n(C.byref(flat.desc()), int(m), float(wl_m)))
