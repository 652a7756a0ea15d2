"""Small geometric helpers user scripts import from optable (reference: optable/solver.py:5-132).

These are host-side conveniences around the traced path (closest approach of two traced rays, the rotation that
takes one normal to another, a vectorised slab test); the device has its own box test (csrc/optb_device.cuh).
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def solve_ray_bboxes_intersections(ray_origin, ray_direction, bboxes):
    """Slab test of one ray against many boxes. `bboxes`: (N, 6) rows (xmin, xmax, ymin, ymax, zmin, zmax).
    Returns (t1, t2, hit): entry/exit parameters clamped to [0, inf) and the hit mask, with the reference's
    conventions: an axis with |d| ~ 0 (np.isclose) only checks containment, a miss on such an axis is reported as
    t1 = 1 > t2 = 0, and hit = (t2 + 1e-12 >= t1) & (t2 >= 0) (solver.py:5-48)."""
    o = np.asarray(ray_origin, dtype=np.float64)
    d = np.asarray(ray_direction, dtype=np.float64)
    boxes = np.asarray(bboxes, dtype=np.float64).reshape(-1, 6)
    n = boxes.shape[0]
    t1, t2 = np.zeros(n), np.full(n, np.inf)
    for ax in range(3):
        lo, hi = boxes[:, 2 * ax], boxes[:, 2 * ax + 1]
        if np.isclose(d[ax], 0.0):
            outside = (o[ax] < lo) | (o[ax] > hi)
            t1[outside], t2[outside] = 1.0, 0.0
            continue
        inv = 1.0 / d[ax]
        ta, tb = (lo - o[ax]) * inv, (hi - o[ax]) * inv
        t1 = np.maximum(t1, np.minimum(ta, tb))
        t2 = np.minimum(t2, np.maximum(ta, tb))
    return t1, t2, (t2 + 1e-12 >= t1) & (t2 >= 0.0)


def solve_ray_ray_intersection(ray1_origin, ray1_direction, ray2_origin, ray2_direction):
    """Closest approach of two rays with both parameters clamped to t >= 0 (solver.py:51-107).
    Returns (t1, t2, P, n): the parameters, the midpoint of the closest-approach segment, and the unit normal of
    the mirror that would turn ray 1 into ray 2 there."""
    p1, p2 = np.array(ray1_origin, dtype=np.float64), np.array(ray2_origin, dtype=np.float64)
    d1, d2 = np.array(ray1_direction, dtype=np.float64), np.array(ray2_direction, dtype=np.float64)
    d1, d2 = d1 / np.linalg.norm(d1), d2 / np.linalg.norm(d2)
    w = p1 - p2
    a, b, c = d1 @ d1, d1 @ d2, d2 @ d2
    dw, ew = d1 @ w, d2 @ w
    det = a * c - b * b
    if det < 1e-6:      # (near-)parallel: keep ray 1 at its origin, project onto ray 2
        t1, t2 = 0.0, ew / c
    else:
        t1, t2 = (b * ew - c * dw) / det, (a * ew - b * dw) / det
    t1, t2 = max(0.0, t1), max(0.0, t2)
    P = 0.5 * ((p1 + t1 * d1) + (p2 + t2 * d2))
    m = -0.5 * (d1 + d2)
    return t1, t2, P, m / np.linalg.norm(m)


def solve_normal_to_normal_rotation(n1, n2) -> Tuple[np.ndarray, float]:
    """Axis and angle (radians) of the rotation taking direction n1 to n2; (x axis, 0) when they are parallel
    to 1e-12 (solver.py:110-132)."""
    u = np.asarray(n1, dtype=np.float64) / np.linalg.norm(n1)
    v = np.asarray(n2, dtype=np.float64) / np.linalg.norm(n2)
    axis = np.cross(u, v)
    s = np.linalg.norm(axis)
    if s < 1e-12:
        return np.array([1, 0, 0]), 0.0
    return axis / s, float(np.arccos(np.clip(u @ v, -1.0, 1.0)))
