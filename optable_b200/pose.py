"""Identity + rigid-pose base classes of the scene description layer.

Mirrors the public surface of the reference's `Base` / `Vector` (optable/base.py:8-156): `_id` inheritance
through `copy()`, every constructor kwarg becomes an attribute, `origin`/`unit`, and the rotation /
translation verbs (`RotX/Y/Z`, `TX/TY/TZ`, `_RotAround`, `Rot?AroundLocal`, `_Translate`). This layer only
*describes* a scene; nothing here is on the per-ray path (that is liboptb.so).
"""
from __future__ import annotations

import copy as _copy

import numpy as np

_AXES = {"x": (1.0, 0.0, 0.0), "y": (0.0, 1.0, 0.0), "z": (0.0, 0.0, 1.0)}


def unit_vector(v) -> np.ndarray:
    v = np.array(v, dtype=float)
    return v / np.linalg.norm(v)


def rotation_matrix(axis, theta: float) -> np.ndarray:
    """Proper rotation by `theta` about `axis` (Rodrigues: I + sin(t) K + (1 - cos(t)) K^2)."""
    ux, uy, uz = unit_vector(axis)
    K = np.array([[0.0, -uz, uy], [uz, 0.0, -ux], [-uy, ux, 0.0]])
    return np.identity(3) + np.sin(theta) * K + (1.0 - np.cos(theta)) * (K @ K)


def rotation_from_x(target) -> np.ndarray:
    """Rotation taking (1,0,0) onto `target` (the frame a Ray reports as transform_matrix)."""
    v = unit_vector(target)
    if np.allclose(v, _AXES["x"]):
        return np.identity(3)
    if np.allclose(v, (-1.0, 0.0, 0.0)):
        return np.diag([-1.0, -1.0, 1.0])
    k = np.cross(_AXES["x"], v)
    return rotation_matrix(k, np.arccos(np.clip(v[0], -1.0, 1.0)))


class Base:
    """Object identity. `copy()` keeps `_id`, so every descendant of a ray shares its family id."""

    def __init__(self, **kwargs):
        if not hasattr(self, "_id"):
            self._id = kwargs.get("id", id(self))
        for key, value in kwargs.items():
            setattr(self, key, value)

    def copy(self, **overrides):
        twin = _copy.deepcopy(self)
        for key, value in overrides.items():
            setattr(twin, key, value)
        return twin


class Vector(Base):
    """Something with a lab-frame position that can be rotated and translated."""

    def __init__(self, origin, **kwargs):
        super().__init__(**kwargs)
        self.origin = np.array(origin, dtype=float)
        self.unit = kwargs.get("unit", 1e-2)  # metres per scene unit (cm by default)

    # helpers kept under the reference's names because user scripts call them
    def _normalize_vector(self, vector) -> np.ndarray:
        return unit_vector(vector)

    def R(self, axis, theta: float) -> np.ndarray:
        return rotation_matrix(axis, theta)

    def _vector_to_R(self, t) -> np.ndarray:
        return rotation_from_x(t)

    # subclasses define how a rotation about a point (given relative to their origin) acts on them
    def _RotAroundLocal(self, axis, localpoint, theta):
        raise NotImplementedError("_RotAroundLocal method not implemented")

    def _RotAroundCenter(self, axis, theta):
        return self._RotAroundLocal(axis, [0, 0, 0], theta)

    def _RotAround(self, axis, point, theta):
        return self._RotAroundLocal(axis, np.array(point) - self.origin, theta)

    def RotX(self, theta):
        return self._RotAroundCenter(_AXES["x"], theta)

    def RotY(self, theta):
        return self._RotAroundCenter(_AXES["y"], theta)

    def RotZ(self, theta):
        return self._RotAroundCenter(_AXES["z"], theta)

    def RotXAroundLocal(self, localpoint, theta):
        return self._RotAroundLocal(_AXES["x"], localpoint, theta)

    def RotYAroundLocal(self, localpoint, theta):
        return self._RotAroundLocal(_AXES["y"], localpoint, theta)

    def RotZAroundLocal(self, localpoint, theta):
        return self._RotAroundLocal(_AXES["z"], localpoint, theta)

    def _Translate(self, movement):
        self.origin += np.array(movement)
        return self

    def TX(self, dx):
        return self._Translate([dx, 0, 0])

    def TY(self, dy):
        return self._Translate([0, dy, 0])

    def TZ(self, dz):
        return self._Translate([0, 0, dz])


def pivot(origin: np.ndarray, R: np.ndarray, localpoint) -> np.ndarray:
    """New origin after rotating by R about the point origin + localpoint."""
    lp = np.array(localpoint, dtype=float)
    return origin + R @ (-lp) + lp
