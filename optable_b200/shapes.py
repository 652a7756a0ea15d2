"""Surface descriptions in a component's local frame (local normal = +x).

Same class names and constructor arguments as the reference's optable/surfaces.py, because user code builds
them directly (`Block(hole=Circle(r))`, `BaseRefraciveSurface(surface=ASphere(...))`, `comp.surface = Polygon(v)`).
Unlike the reference these objects carry parameters only: the per-ray arithmetic (f, normal, within_boundary,
root finding) lives in csrc/optb_device.cuh. What the scene flattener reads from each class:

  Circle.radius | Rectangle.width/.height | Sphere.radius/.height/.diameter | Cylinder.radius/.height/.theta_range
  ASphere.radius/.xmin/.xmax/.asphere_spec | Polygon.vertices/._normal/._basis/._verts2d/._bbox/.planar
  Plane.csg = (op, A, B) for union / subtract | get_bbox_local() everywhere
"""
from __future__ import annotations

from typing import Sequence

import numpy as np

from .pose import Base, unit_vector


def _merge(b1, b2):
    return tuple(f(b1[i], b2[i]) for i, f in enumerate((min, max, min, max, min, max)))


class Surface(Base):
    planar = True

    def __init__(self):
        super().__init__()
        self.planar = type(self).planar

    def get_bbox_local(self):
        raise NotImplementedError("Method 'get_bbox_local' must be implemented in the derived class.")

    @staticmethod
    def merge_bbox(bbox1, bbox2):
        return _merge(bbox1, bbox2)

    @staticmethod
    def merge_bboxs(bboxs):
        cols = list(zip(*bboxs))
        return tuple((np.min if i % 2 == 0 else np.max)(cols[i]) for i in range(6))


class Point(Surface):
    """Degenerate marker; rays never hit it (the reference's f = |P| has no sign change)."""
    planar = False

    def get_bbox_local(self):
        return (0, 0, 0, 0, 0, 0)


class Plane(Surface):
    """The x = 0 plane. A bare Plane has no boundary; bounded apertures derive from it, and
    `union` / `subtract` compose two of them (surfaces.py:100-136)."""

    def __init__(self):
        super().__init__()
        self._normal = np.array([1, 0, 0])
        self.csg = None

    def _compose(self, op, other):
        out = Plane()
        out.csg = (op, self, other)
        return out

    def union(self, other):
        return self._compose("union", other)

    def subtract(self, other):
        return self._compose("subtract", other)

    def get_bbox_local(self):
        if self.csg is None:
            raise NotImplementedError("Method 'get_bbox_local' must be implemented in the derived class.")
        return _merge(self.csg[1].get_bbox_local(), self.csg[2].get_bbox_local())


class Circle(Plane):
    def __init__(self, radius):
        super().__init__()
        self.radius = radius

    def get_bbox_local(self):
        r = self.radius
        return (0, 0, -r, r, -r, r)


class Rectangle(Plane):
    def __init__(self, width, height):
        super().__init__()
        self.width = width    # along local y
        self.height = height  # along local z

    def get_bbox_local(self):
        return (0, 0, -self.width / 2, self.width / 2, -self.height / 2, self.height / 2)


class Cylinder(Surface):
    """Cylinder about the local z axis, limited to theta_range (atan2(y, x)) and |z| <= height/2."""
    planar = False

    def __init__(self, radius, height, theta_range=(-np.pi, np.pi)):
        super().__init__()
        self.radius, self.height, self.theta_range = radius, height, theta_range

    def get_bbox_local(self):
        r, hh = self.radius, self.height / 2
        return (-r, r, -r, r, -hh, hh)


class Sphere(Surface):
    """Spherical cap of a sphere centred on the local origin: radius - height <= x <= radius."""
    planar = False

    def __init__(self, radius, height=None):
        super().__init__()
        self.radius = radius
        self.height = height if height is not None else 2 * radius
        # the reference evaluates this with the raw argument, so height=None is a TypeError there too
        self.diameter = np.sqrt(radius ** 2 - (radius - height) ** 2) * 2

    def get_bbox_local(self):
        half = self.diameter / 2
        return (self.radius - self.height, self.radius, -half, half, -half, half)


class ASphere(Surface):
    """Rotationally symmetric profile x = -f_asphere(r), r = |(y, z)| <= radius.

    The device needs the profile in closed form. `asphere_spec = (form, params)` names it; callables built by
    this package's lens classes carry it as an attribute, and the two closure shapes of the reference's lens
    classes are recognised by the flattener. Any other callable is rejected at flatten time (no CPU path).
    """
    planar = False

    def __init__(self, radius, f_asphere: callable):
        super().__init__()
        self.radius = radius
        self.f_asphere = f_asphere
        self.asphere_spec = getattr(f_asphere, "asphere_spec", None)
        x0, xR = -f_asphere(0), -f_asphere(radius)
        self.xmin, self.xmax = min(x0, xR), max(x0, xR)

    def get_bbox_local(self):
        r = self.radius
        return (self.xmin, self.xmax, -r, r, -r, r)

    def roc_r(self, r: float) -> float:
        """Local radius of curvature (1 + f'^2)^1.5 / f'' from central differences with h = 1e-4 radius, the
        stencil the reference uses (surfaces.py:351-369); the device repeats it per hit. Its presence makes a
        refractive component treat `roc` as position dependent."""
        h = 1e-4 * self.radius
        f = self.f_asphere
        d1 = (f(r + h) - f(r - h)) / (2 * h)
        d2 = (f(r + h) - 2 * f(r) + f(r - h)) / (h ** 2)
        return (1 + d1 ** 2) ** 1.5 / d2

    def roc(self, P) -> float:
        return self.roc_r(float(np.hypot(P[1], P[2])))


def parametric_asphere(R, kappa, a4=0, a6=0, a8=0):
    """Conic + even polynomial sag (component_group.py:1096-1101)."""

    def sag(r):
        conic = r ** 2 / (R * (1 + np.sqrt(1 - (1 + kappa) * (r ** 2) / (R ** 2))))
        return conic + a4 * r ** 4 + a6 * r ** 6 + a8 * r ** 8

    sag.asphere_spec = ("parametric", dict(R=R, kappa=kappa, a4=a4, a6=a6, a8=a8))
    return sag


def exact_spherical_asphere(EFL, n):
    """Aberration-free plano-convex hyperboloid (component_group.py:1067-1070)."""

    def sag(r):
        return (EFL / (n + 1)) * (-1 + np.sqrt(1 + (n + 1) / (n - 1) * (r ** 2) / (EFL ** 2)))

    sag.asphere_spec = ("exact_spherical", dict(EFL=EFL, n=n))
    return sag


class Polygon(Plane):
    """Flat polygon. (N,2) vertices, or (N,3) with x = 0, lie in the local x = 0 plane (planar=True); any
    other coplanar (N,3) loop is a tilted facet handled by the curved-surface branch (planar=False).
    Vertices counter-clockwise seen along the normal (surfaces.py:426-516)."""

    def __init__(self, vertices: Sequence[Sequence[float]], normal=None):
        super().__init__()
        self._tol = 1e-9
        verts = np.asarray(vertices, dtype=float)
        if verts.ndim != 2 or verts.shape[0] < 3:
            raise ValueError("Need at least three vertices (shape (N,2) or (N,3)).")
        self.planar = False
        if verts.shape[1] == 2:
            verts = np.column_stack((np.zeros(len(verts)), verts))
            if normal is None:
                normal, self.planar = (1.0, 0.0, 0.0), True
        if verts.shape[1] == 3 and normal is None and np.allclose(verts[:, 0], 0.0):
            normal, self.planar = (1.0, 0.0, 0.0), True
        self.vertices = verts
        if normal is None:
            for i in range(2, len(verts)):
                candidate = np.cross(verts[i] - verts[0], verts[1] - verts[0])
                if np.linalg.norm(candidate) > self._tol:
                    normal = candidate
                    break
            else:
                raise ValueError("Vertices are colinear - cannot define a plane.")
        self._normal = unit_vector(normal)
        if np.any(np.abs((verts - verts[0]) @ self._normal) > self._tol):
            raise ValueError("Vertices are not coplanar with the supplied normal.")
        helper = np.array([1.0, 0.0, 0.0])
        if abs(float(helper @ self._normal)) > 0.99:
            helper = np.array([0.0, 1.0, 0.0])
        u = unit_vector(np.cross(self._normal, helper))
        self._basis = (u, np.cross(self._normal, u))
        rel = verts - verts[0]
        self._verts2d = np.column_stack((rel @ self._basis[0], rel @ self._basis[1]))
        lo, hi = verts.min(axis=0), verts.max(axis=0)
        self._bbox = (lo[0], hi[0], lo[1], hi[1], lo[2], hi[2])

    def get_bbox_local(self):
        return self._bbox
