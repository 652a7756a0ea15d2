"""Parity comparison between two result sets in the layout of oracle.ref_harness.run_reference.

Bar (BASELINE.json north_star): hit component index and bounce count bit-exact; positions, directions,
path lengths and q-parameters within 1e-9 relative in fp64. "Relative" for a vector is taken against the
vector's norm (floored at 1 scene unit for positions, so a coordinate that is exactly 0 in one result and
1e-17 in the other is not a failure).
"""
from __future__ import annotations

import numpy as np

RTOL = 1e-9


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
        errs["seg_length"] = _rel(ref["seg_length"], got["seg_length"], 1e-3).max()
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
        errs["hit_t"] = _rel(ref["hit_t"], got["hit_t"], 1e-3).max()
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
    itself. Scenes with such surfaces compare q to 1e-6; every other field and scene keeps the 1e-9 bar."""
    from optable_b200 import _abi as A

    return 1e-6 if (flat.node_i[:, A.NI_ROCKIND] == A.ROC_ASPHERE_FD).any() else RTOL
