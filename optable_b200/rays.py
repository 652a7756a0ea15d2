"""Ray state and Gaussian-beam q algebra (reference: optable/ray.py:8-216, 428-445).

A `Ray` is the host-side record of one ray or one traced segment; the device carries the same fields as SoA
fp64 arrays (include/optb.h optb_rays). `RayBundle` is the tensor-native container for batches too large to
exist as Python objects.
"""
from __future__ import annotations

from typing import List

import numpy as np

from .materials import RefractiveIndex
from .pose import Vector, pivot, rotation_from_x, unit_vector


class GaussianBeam:
    """Complex beam parameter q = z + i z_R helpers (ray.py:8-55)."""

    @staticmethod
    def q_at_waist(w0, wl, n=1):
        return (1j * n * np.pi * w0 ** 2) / wl

    @staticmethod
    def q_at_z(qo, z):
        return qo + z

    @staticmethod
    def distance_to_waist(q):
        return np.real(q)

    @staticmethod
    def rayleigh_range(q):
        return np.imag(q)

    @staticmethod
    def waist(q, wl, n=1):
        return np.sqrt((wl * np.imag(q)) / (n * np.pi))

    @staticmethod
    def radius_of_curvature(q):
        return 1 / np.real(1 / q)

    @staticmethod
    def spot_size(qo, z, wl, n=1):
        return np.sqrt(-wl / (n * np.pi * np.imag(1 / (qo + z))))


class Ray(Vector):
    """Geometric ray with intensity, wavelength, optional Gaussian q at its origin and accumulated path length."""

    _n = RefractiveIndex("_n")

    def __init__(self, origin, direction, intensity: float = 1.0, wavelength=None, length=None, alive=True,
                 qo=None, w0=None, **kwargs):
        super().__init__(origin, **kwargs)
        self.length = float(length) if length else None
        self.direction = direction
        self.intensity = float(intensity)
        self.wavelength = float(wavelength) if wavelength else 0.0
        self.alive = alive
        self._n = 1.0
        self._pathlength = 0.0
        if qo is not None:
            self.qo = qo
        elif w0 is not None:
            self.qo = self.q_at_waist(w0)
        else:
            self.qo = None

    def __repr__(self):
        return (f"Ray(origin={self.origin}, direction={self.direction}, intensity={self.intensity}, "
                f"length={self.length}, alive={self.alive}, qo={self.qo})")

    def copy(self, **overrides):
        """Same `_id`, fresh arrays. (The generic deepcopy is not needed: a Ray owns only two small arrays.)"""
        twin = object.__new__(type(self))
        fields = twin.__dict__
        fields.update(self.__dict__)
        fields["origin"] = np.array(self.origin, dtype=float)
        fields["_direction"] = np.array(self._direction, dtype=float)
        for key, value in overrides.items():
            setattr(twin, key, value)
        return twin

    @property
    def direction(self) -> np.ndarray:
        return self._direction

    @direction.setter
    def direction(self, value):
        self._direction = unit_vector(value)

    @property
    def n(self) -> float:
        return self._n()

    @property
    def transform_matrix(self) -> np.ndarray:
        return rotation_from_x(self.direction)

    @property
    def tangent_1(self) -> np.ndarray:
        d = self.direction
        if d[0] == 0 and d[1] == 0:
            return np.array([1, 0, 0])
        return unit_vector(np.cross(d, [0, 0, 1]))

    @property
    def tangent_2(self) -> np.ndarray:
        return unit_vector(np.cross(self.direction, self.tangent_1))

    def pathlength(self, t: float = 0) -> float:
        return float(self._pathlength + t * self.n)

    def phase(self, t: float = 0) -> float:
        return np.mod((2 * np.pi / self.wavelength) * self.pathlength(t), 2 * np.pi)

    def _RotAroundLocal(self, axis, localpoint, theta) -> "Ray":
        R = self.R(axis, theta)
        self.direction = R @ self.direction
        self.origin = pivot(self.origin, R, localpoint)
        return self

    # Gaussian beam conveniences bound to this ray's wavelength / index
    def q_at_waist(self, w0):
        return GaussianBeam.q_at_waist(w0, self.wavelength, self.n)

    def q_at_z(self, z):
        return GaussianBeam.q_at_z(self.qo, z)

    def distance_to_waist(self, q):
        return GaussianBeam.distance_to_waist(q)

    def waist(self, q):
        return GaussianBeam.waist(q, self.wavelength, self.n)

    def rayleigh_range(self, q):
        return GaussianBeam.rayleigh_range(q)

    def radius_of_curvature(self, q):
        return GaussianBeam.radius_of_curvature(q)

    def spot_size(self, z):
        return GaussianBeam.spot_size(self.qo, z, self.wavelength, self.n)

    def Propagate(self, z) -> "Ray":
        return self.copy(qo=self.q_at_z(z))


def multiplex_rays_in_wavelength(rays: List[Ray], wavelength_list: List[float]) -> List[Ray]:
    """One copy of every ray per wavelength, wavelength-major; copies keep the source `_id` (ray.py:428-445)."""
    return [ray.copy(wavelength=wl) for wl in wavelength_list for ray in rays]
