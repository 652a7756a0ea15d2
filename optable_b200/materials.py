"""Optical materials: constant and Sellmeier-3 refractive indices (reference: optable/material.py).

`Material.n(wavelength_m)` takes metres. `RefractiveIndex` is the descriptor the reference uses to bind a
material to the wavelength of the object that carries it (material.py:48-85): reading `obj._n` yields a
callable that evaluates the stored material at `obj.wavelength * obj.unit`.
The device evaluates the same formulas (csrc/optb_device.cuh material_n); these host versions serve scene
construction (e.g. deriving a lens radius from an index) and the public API.
"""
from __future__ import annotations

from typing import Callable, Sequence, Union

import numpy as np


class Material:
    def __init__(self, name: str, n: Union[Callable, float]):
        self.name = name
        if isinstance(n, (int, float)):
            self.n_const = float(n)          # read by the scene flattener
            self.n_func = self._constant
        else:
            self.n_const = None
            self.n_func = n

    def _constant(self, wavelength_m):
        return self.n_const

    def n(self, wavelength_m: float) -> float:
        """Refractive index at a wavelength given in metres."""
        return self.n_func(wavelength_m)


class ConstMaterial(Material):
    def __init__(self, name: str = "", n: float = 1.0):
        super().__init__(name, n)


class Vacuum(Material):
    def __init__(self):
        super().__init__("Vacuum", n=1.0)


class SellmeierMaterial(Material):
    """n^2 = 1 + sum_i B_i L^2 / (L^2 - C_i), L in micrometres, C_i in um^2 (material.py:106-120)."""

    def __init__(self, name: str, Bs: Sequence[float], Cs: Sequence[float]):
        self.Bs = Bs
        self.Cs = Cs
        super().__init__(name, self.sellmeier_n)

    def sellmeier_n(self, wavelength_m):
        lam2 = (wavelength_m / 1e-6) ** 2
        total = 1.0
        for B, Cc in zip(self.Bs, self.Cs):
            total += B * lam2 / (lam2 - Cc)
        return np.sqrt(total)


def _glass(name, Bs, Cs):
    def __init__(self):
        SellmeierMaterial.__init__(self, name, list(Bs), list(Cs))

    return __init__


class Glass_NBK7(SellmeierMaterial):
    __init__ = _glass("BK7", (1.03961212, 0.231792344, 1.01046945), (0.00600069867, 0.0200179144, 103.560653))


class Glass_UVFS(SellmeierMaterial):
    __init__ = _glass("UV Fused Silica", (0.6961663, 0.4079426, 0.8974794),
                      (0.0684043 ** 2, 0.1162414 ** 2, 9.896161 ** 2))


class Glass_NSF5(SellmeierMaterial):
    __init__ = _glass("N_SF5", (1.52481889, 0.187085527, 1.42729015), (0.011254756, 0.0588995392, 129.141675))


class Glass_NSF11(SellmeierMaterial):
    __init__ = _glass("N_SF11", (1.73759695, 0.313747346, 1.89878101), (0.013188707, 0.0623068142, 155.23629))


class Glass_NSK2(SellmeierMaterial):
    __init__ = _glass("N_SK2", (1.28189012, 0.257738258, 0.96818604), (0.0072719164, 0.0242823527, 110.377773))


class Glass_NSF57(SellmeierMaterial):
    __init__ = _glass("N_SF57", (1.87543481, 0.37375749, 2.30001797), (0.0141749518, 0.0640509927, 177.389795))


class RefractiveIndex:
    """Descriptor: stores a Material in the instance dict, reads back as n(wavelength_m=None) -> float."""

    def __init__(self, storage_name: str):
        self.storage_name = storage_name

    def __set__(self, instance, value):
        if not isinstance(value, Material):
            value = Material("Constant", n=float(value))
        instance.__dict__[self.storage_name] = value

    def __get__(self, instance, owner):
        if instance is None:
            return self
        material = instance.__dict__.get(self.storage_name)
        if material is None:
            raise AttributeError(f"Material for {self.storage_name} not initialized.")

        def evaluate(wavelength_m=None):
            if wavelength_m is None:
                wl = getattr(instance, "wavelength", None)
                wavelength_m = wl * instance.unit if wl is not None else 0.0
            return float(material.n(wavelength_m))

        return evaluate
