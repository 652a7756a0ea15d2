"""Parity comparison between two result sets in the layout of oracle.ref_harness.run_reference.

Bar (BASELINE.json north_star): hit component index and bounce count bit-exact; positions, directions,
path lengths and q-parameters within 1e-9 relative in fp64. "Relative" for a vector is taken against the
vector's norm (floored at 1 scene unit for positions, so a coordinate that is exactly 0 in one result and
1e-17 in the other is not a failure).
"""
from __future__ import annotations

import numpy as np

RTOL = 1e-9
Q_RTOL_FD = 2e-7   # Gaussian q behind a surface whose curvature the reference takes from a finite-difference stencil
# Distances along a ray are compared relative to max(|t|, LENGTH_FLOOR): the reference takes a curved-surface hit from
# scipy's brentq with xtol = 2e-12 (absolute), so its own t is only defined to 2e-12; RTOL * LENGTH_FLOOR is that.
LENGTH_FLOOR = 2e-3


def _rel_vec(a, b, floor):
    scale = np.maximum(np.linalg.norm(a, axis=-1), floor)
    return np.linalg.norm(a - b, axis=-1) / scale


def _rel(a, b, floor=0.0):
    a, b = np.asarray(a), np.asarray(b)
    both_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
    scale = np.maximum(np.abs(a), floor)
    with np.errstate(invalid="ignore", divide="ignore"):
        r = np.abs(a - b) / np.where(scale == 0, 1.0, scale)
    r = np.where(both_inf, 0.0, r)
    r = np.where((a == b), 0.0, r)
    return r


def compare(ref, got, rtol=RTOL, q_rtol=None, label=""):
    """Raise AssertionError with a readable message on the first mismatch; return max relative errors."""
    q_rtol = rtol if q_rtol is None else q_rtol
    errs = {}
    assert len(ref["seg_root"]) == len(got["seg_root"]), \
        f"{label}: segment count {len(got['seg_root'])} != reference {len(ref['seg_root'])}"
    for k in ("seg_root", "seg_pop", "seg_leaf", "seg_alive", "seg_hasq"):
        bad = np.nonzero(np.asarray(ref[k]) != np.asarray(got[k]))[0]
        assert bad.size == 0, f"{label}: {k} differs at rows {bad[:8]} (ref {np.asarray(ref[k])[bad[:8]]}, got {np.asarray(got[k])[bad[:8]]})"
    if len(ref["seg_root"]):
        errs["seg_o"] = _rel_vec(ref["seg_o"], got["seg_o"], 1.0).max()
        errs["seg_d"] = _rel_vec(ref["seg_d"], got["seg_d"], 1.0).max()
        errs["seg_length"] = _rel(ref["seg_length"], got["seg_length"], LENGTH_FLOOR).max()
        errs["seg_intensity"] = _rel(ref["seg_intensity"], got["seg_intensity"]).max()
        errs["seg_wavelength"] = _rel(ref["seg_wavelength"], got["seg_wavelength"]).max()
        errs["seg_pathlength"] = _rel(ref["seg_pathlength"], got["seg_pathlength"], 1e-3).max()
        errs["seg_n"] = _rel(ref["seg_n"], got["seg_n"]).max()
        hq = np.asarray(ref["seg_hasq"])
        if hq.any():
            qa, qb = np.asarray(ref["seg_q"])[hq], np.asarray(got["seg_q"])[hq]
            errs["seg_q"] = (np.abs(qa - qb) / np.maximum(np.abs(qa), 1e-300)).max()
    assert len(ref["hit_root"]) == len(got["hit_root"]), \
        f"{label}: monitor hit count {len(got['hit_root'])} != reference {len(ref['hit_root'])}"
    for k in ("hit_monitor", "hit_root", "hit_pop"):
        bad = np.nonzero(np.asarray(ref[k]) != np.asarray(got[k]))[0]
        assert bad.size == 0, f"{label}: {k} differs at rows {bad[:8]}"
    if len(ref["hit_root"]):
        errs["hit_P"] = _rel_vec(ref["hit_P"], got["hit_P"], 1.0).max()
        errs["hit_t"] = _rel(ref["hit_t"], got["hit_t"], LENGTH_FLOOR).max()
        errs["hit_intensity"] = _rel(ref["hit_intensity"], got["hit_intensity"]).max()
        errs["hit_d"] = _rel_vec(ref["hit_d"], got["hit_d"], 1.0).max()
    for k, v in errs.items():
        tol = q_rtol if k == "seg_q" else rtol
        assert v <= tol, f"{label}: {k} max relative error {v:.3e} > {tol:.1e}"
    return errs


def sort_hits_reference_order(arrs):
    """Reorder hit rows to (root, monitor, pop): the order Monitor.record produces when the reference
    traces the initial rays one after another."""
    key = np.lexsort((arrs["hit_pop"], arrs["hit_monitor"], arrs["hit_root"]))
    return {k: (v[key] if k.startswith("hit_") else v) for k, v in arrs.items()}


def q_rtol_for(flat):
    """Tolerance on the Gaussian q for a scene. After an ASphere, q depends on a radius of curvature that the
    reference takes from a finite-difference second derivative (surfaces.py:355-369, h = 1e-4 radius): one ulp in
    the local hit point or in f_asphere moves ROC by ~5e-9 relative (rounding noise / h^2), in the reference
    itself: the curvature the REFERENCE computes is a smooth function of the hit point plus rounding noise of ~1e-8
    relative that is resampled by any change of the hit point, however small. Scenes with such surfaces compare q to
    2e-7 (SURVEY A.11 budgeted ~1e-7 per surface; observed 3e-8 at 4096 rays, 7e-8 at 3e6 rays through two lenses);
    every other field and scene keeps the 1e-9 bar."""
    from optable_b200 import _abi as A

    return Q_RTOL_FD if (flat.node_i[:, A.NI_ROCKIND] == A.ROC_ASPHERE_FD).any() else RTOL


def _without_roots(arrs, roots):
    keep_s = ~np.isin(arrs["seg_root"], roots)
    keep_h = ~np.isin(arrs["hit_root"], roots)
    return {k: (np.asarray(v)[keep_s] if k.startswith("seg_") else np.asarray(v)[keep_h] if k.startswith("hit_") else v)
            for k, v in arrs.items()}


def compare_flagging_ties(flat, ref, got, rtol=RTOL, q_rtol=None, label="", tie_rtol=1e-9):
    """`compare`, with the exclusion BASELINE.json's north_star states for grazing/edge cases made explicit.

    An initial ray is set aside (flagged) when its two traces part ways at a pop where the two competing surfaces
    are hit at the same distance to within `tie_rtol`: coplanar overlapping apertures (lenslets of a micro-lens
    array with radius > pitch/2), or a ray through an edge shared by two faces. Which of two equal distances is
    the smaller one is decided by the last bit of t -- in the reference as well -- and everything after that pop
    legitimately differs. Any other difference fails as in `compare`. Returns (errs, sorted list of flagged roots);
    callers bound the flagged fraction."""
    from oracle import oracle as O
    from optable_b200 import _abi as A

    node_of_leaf = {int(flat.node_i[i, A.NI_LEAF]): i for i in range(flat.n_nodes) if flat.node_i[i, A.NI_LEAF] >= 0}
    flagged = []
    for root in np.union1d(ref["seg_root"], got["seg_root"]):
        a, b = np.nonzero(ref["seg_root"] == root)[0], np.nonzero(got["seg_root"] == root)[0]
        m = min(len(a), len(b))
        diff = np.nonzero((ref["seg_pop"][a[:m]] != got["seg_pop"][b[:m]]) | (ref["seg_leaf"][a[:m]] != got["seg_leaf"][b[:m]]))[0]
        if diff.size == 0 and len(a) == len(b):
            continue
        assert diff.size, f"{label}: root {root} has {len(b)} segments, reference {len(a)}, with an identical common prefix"
        ka, kb = a[diff[0]], b[diff[0]]
        la, lb = int(ref["seg_leaf"][ka]), int(got["seg_leaf"][kb])
        same_ray = (ref["seg_pop"][ka] == got["seg_pop"][kb] and _rel_vec(ref["seg_o"][ka], got["seg_o"][kb], 1.0) <= rtol
                    and _rel_vec(ref["seg_d"][ka], got["seg_d"][kb], 1.0) <= rtol)
        assert same_ray and la >= 0 and lb >= 0, \
            f"{label}: root {root} pop {ref['seg_pop'][ka]}: leaf {lb} != reference {la} and the popped rays differ or one side missed"
        ts = [O.intersect(flat, node_of_leaf[l], ref["seg_o"][ka], ref["seg_d"][ka])[1] for l in (la, lb)]
        assert all(t is not None and np.isfinite(t) for t in ts) and abs(ts[0] - ts[1]) <= tie_rtol * max(abs(ts[0]), 1e-3), \
            f"{label}: root {root} pop {ref['seg_pop'][ka]}: leaf {lb} (t={ts[1]!r}) != reference {la} (t={ts[0]!r}): not a tie"
        flagged.append(int(root))
    if flagged:
        ref, got = _without_roots(ref, flagged), _without_roots(got, flagged)
    return compare(ref, got, rtol=rtol, q_rtol=q_rtol, label=label), flagged


def restart_batch(raw, roots_without_length_limit):
    """Every popped ray of a finished trace (raw result arrays of oracle.trace / Engine.trace_arrays) as a fresh
    batch of initial rays. Tracing this batch for a few pops compares single interactions on IDENTICAL inputs, so
    the comparison is free of the error growth along long chaotic paths (a curved-surface hit is only defined to
    brentq's xtol = 2e-12 in the reference, and trapped rays amplify that by a factor per bounce)."""
    from optable_b200 import _abi as A

    keep = np.isin(raw["seg_root"], roots_without_length_limit)
    n = int(keep.sum())
    out = {"ox": raw["seg_ox"][keep], "oy": raw["seg_oy"][keep], "oz": raw["seg_oz"][keep],
           "dx": raw["seg_dx"][keep], "dy": raw["seg_dy"][keep], "dz": raw["seg_dz"][keep],
           "intensity": raw["seg_intensity"][keep], "wavelength": raw["seg_wavelength"][keep],
           "q_re": raw["seg_q_re"][keep], "q_im": raw["seg_q_im"][keep], "pathlength": raw["seg_pathlength"][keep],
           "n_medium": raw["seg_n"][keep], "length": np.full(n, np.inf)}
    out = {k: np.ascontiguousarray(v, dtype=np.float64) for k, v in out.items()}
    out["flags"] = (np.uint32(A.RF_ALIVE) | (raw["seg_flags"][keep] & np.uint32(A.RF_HASQ))).astype(np.uint32)
    out["family"] = np.arange(n, dtype=np.int32)
    return out
