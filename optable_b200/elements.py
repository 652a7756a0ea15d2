"""Leaf optical components: a pose (origin + local->lab rotation), one surface and the interaction parameters.

Same names, constructor arguments and attributes as the reference's leaf classes
(optable/optical_component.py:8-124, 429-974); the class hierarchy (BaseMirror / BaseRefraciveSurface / Lens /
Block / PointObj) is what the scene flattener keys the device interaction kind on. The per-ray methods of the
reference (`interact`, `intersect_point_local`, `interact_local`) are NOT reimplemented here: that arithmetic
runs in liboptb.so (csrc/optb_device.cuh), and there is deliberately no host fallback.
"""
from __future__ import annotations

import numpy as np

from .materials import RefractiveIndex
from .pose import Vector, pivot
from .shapes import Circle, Cylinder, Plane, Point, Rectangle, Sphere

_UNSET = (None,) * 6


class OpticalComponent(Vector):
    def __init__(self, origin, **kwargs):
        super().__init__(origin, **kwargs)
        self.transform_matrix = np.identity(3)
        self.surface = Plane()
        self._bbox = _UNSET
        self.render_obj = kwargs.get("render_obj", True)
        self.render_comp_vec = kwargs.get("render_comp_vec", False)
        self.name = kwargs.get("name", None)
        self.label = kwargs.get("label", None)
        self.label_position = kwargs.get("label_position", [1, 0, 0])
        self._interact_count = {}
        self.max_interact_count = kwargs.get("max_interact_count", None)

    def __repr__(self):
        return f"{type(self).__name__}(origin={self.origin}, transform_matrix=\n{self.transform_matrix})"

    # local axes in the lab frame
    @property
    def normal(self):
        return self.transform_matrix @ np.array([1, 0, 0])

    @property
    def tangent_Y(self):
        return self.transform_matrix @ np.array([0, 1, 0])

    @property
    def tangent_Z(self):
        return self.transform_matrix @ np.array([0, 0, 1])

    def get_bbox_local(self):
        return self.surface.get_bbox_local()

    @property
    def bbox(self):
        """Lab AABB, computed on first use and then kept (the reference never invalidates it either:
        optical_component.py:62-67), so a component moved after its first trace keeps its old box."""
        if self._bbox == _UNSET:
            self._bbox = tuple(self.get_bbox())
        return tuple(self._bbox)

    def get_bbox(self) -> tuple:
        x0, x1, y0, y1, z0, z1 = self.get_bbox_local()
        corners = np.array([[x, y, z] for z in (z0, z1) for y in (y0, y1) for x in (x0, x1)], dtype=float).T
        lab = self.transform_matrix @ corners + self.origin.reshape(3, 1)
        lo, hi = lab.min(axis=1), lab.max(axis=1)
        return (lo[0], hi[0], lo[1], hi[1], lo[2], hi[2])

    def _RotAroundLocal(self, axis, localpoint, theta):
        R = self.R(axis, theta)
        self.transform_matrix = R @ self.transform_matrix
        self.origin = pivot(self.origin, R, localpoint)
        return self

    def interact(self, ray, engine=None):
        """One pop of the bounce loop against this component (or group) alone, on the device: (t, [truncated
        parent, children...]) or (None, None), like optical_component.py:337-378 / component_group.py:93-122."""
        from .table import single_pop

        return single_pop(self, ray, engine)

    def gather_components(self, avoid_flatten_classname=(), ignore_classname=()):
        """Flat description rows of this component tree for the CSV export (optical_component.py:386-426)."""
        from .export import component_rows

        return component_rows(self, avoid_flatten_classname, ignore_classname)

    def point_to_lab_coordinates(self, point_local):
        return self.transform_matrix @ np.asarray(point_local, dtype=float) + self.origin

    # interact-count bookkeeping (optical_component.py:136-149); the device keeps the live table
    def get_interact_count(self, ray_id):
        return self._interact_count.get(ray_id, 0)

    def should_interact(self, ray_id):
        return self.max_interact_count is None or self.get_interact_count(ray_id) < self.max_interact_count

    def increase_interact_count(self, ray_id):
        self._interact_count[ray_id] = self.get_interact_count(ray_id) + 1

    def patch_block(self, width, height):
        """Opaque frame of width x height around this component's aperture, sharing its pose."""
        frame = Block(self.origin, hole=self.surface, width=width, height=height)
        frame.transform_matrix = self.transform_matrix
        return frame


class PointObj(OpticalComponent):
    """Reference marker; never intercepts rays."""

    def __init__(self, origin, **kwargs):
        super().__init__(origin, **kwargs)
        self.surface = Point()


class Block(OpticalComponent):
    """Absorbing rectangle, optionally with a hole cut out."""

    def __init__(self, origin, hole=None, width: float = 1.0, height: float = 1.0, **kwargs):
        super().__init__(origin, **kwargs)
        self.width, self.height = width, height
        plate = Rectangle(width, height)
        self.surface = plate if hole is None else plate.subtract(hole)


class BaseMirror(OpticalComponent):
    """Reflecting surface: a reflected child when reflectivity > 0 and a straight-through child when
    transmission > 0, in that order."""

    def __init__(self, origin, reflectivity: float = 1.0, transmission: float = 0.0, **kwargs):
        super().__init__(origin, **kwargs)
        self.reflectivity, self.transmission = reflectivity, transmission


class BaseRefraciveSurface(OpticalComponent):
    """Interface between index n1 (local x > 0 side) and n2 (x < 0 side): Snell refraction or total internal
    reflection, plus an extra reflected child when reflectivity > 0. (Spelling follows the reference.)"""

    _n1 = RefractiveIndex("_n1")
    _n2 = RefractiveIndex("_n2")

    def __init__(self, origin, n1=1.0, n2=1.0, reflectivity: float = 0.0, transmission: float = 1.0, **kwargs):
        super().__init__(origin, **kwargs)
        self._n1, self._n2 = n1, n2
        self.reflectivity, self.transmission = reflectivity, transmission
        self.surface = kwargs.get("surface", Plane())
        self.roc = self.surface.roc if hasattr(self.surface, "roc") else np.inf


class Mirror(BaseMirror):
    def __init__(self, origin, radius: float = 0.5, reflectivity: float = 1.0, transmission: float = 0.0, **kwargs):
        super().__init__(origin, reflectivity=reflectivity, transmission=transmission, **kwargs)
        self.radius = radius
        self.surface = Circle(radius)


class SquareMirror(BaseMirror):
    def __init__(self, origin, width: float = 1.0, height: float = 1.0, reflectivity: float = 1.0,
                 transmission: float = 0.0, **kwargs):
        super().__init__(origin, reflectivity=reflectivity, transmission=transmission, **kwargs)
        self.width, self.height = width, height
        self.surface = Rectangle(width, height)


class BeamSplitter(SquareMirror):
    """Partially reflecting plate: amplitude-like split sqrt(eta) / sqrt(1 - eta)."""

    def __init__(self, origin, width=1.0, height=1.0, eta: float = 0.5, **kwargs):
        super().__init__(origin, width=width, height=height, reflectivity=np.sqrt(eta),
                         transmission=np.sqrt(1 - eta), **kwargs)


class CylMirror(BaseMirror):
    def __init__(self, origin, radius: float = 0.5, height: float = 1.0, theta_range=(-np.pi, np.pi), **kwargs):
        super().__init__(origin, **kwargs)
        self.radius, self.height = radius, height
        self.surface = Cylinder(radius, height, theta_range)


class SquareRefractive(BaseRefraciveSurface):
    def __init__(self, origin, width: float = 1.0, height: float = 1.0, n1=1.0, n2=1.0, reflectivity: float = 0.0,
                 transmission: float = 1.0, **kwargs):
        super().__init__(origin, n1=n1, n2=n2, reflectivity=reflectivity, transmission=transmission, **kwargs)
        self.width, self.height = width, height
        self.surface = Rectangle(width, height)


class CircleRefractive(BaseRefraciveSurface):
    def __init__(self, origin, radius: float = 0.5, n1=1.0, n2=1.0, reflectivity: float = 0.0,
                 transmission: float = 1.0, **kwargs):
        super().__init__(origin, n1=n1, n2=n2, reflectivity=reflectivity, transmission=transmission, **kwargs)
        self.radius = radius
        self.surface = Circle(radius)


class SphereRefractive(BaseRefraciveSurface):
    """Spherical cap; the component origin is the centre of curvature and roc = +radius."""

    def __init__(self, origin, radius: float = 0.5, height: float = 0.5, n1=1.0, n2=1.0, reflectivity: float = 0.0,
                 transmission: float = 1.0, **kwargs):
        super().__init__(origin, n1=n1, n2=n2, reflectivity=reflectivity, transmission=transmission, **kwargs)
        self.radius, self.height = radius, height
        self.roc = radius
        self.surface = Sphere(radius, height)


class Lens(OpticalComponent):
    """Ideal thin lens with a circular aperture."""

    def __init__(self, origin, focal_length: float, radius: float = 0.5, transmission: float = 1.0, **kwargs):
        super().__init__(origin, **kwargs)
        self.focal_length, self.transmission = focal_length, transmission
        self.radius = radius
        self.surface = Circle(radius)
