"""Rigid assemblies of leaf components (reference: optable/component_group.py).

`ComponentGroup` keeps a list of children that rotate and translate together and caches the merged lab AABB
that the closest-hit search uses to cull the whole subtree (component_group.py:28-47, 93-122; on the device:
node kind OPTB_G_GROUP + skip pointer). The concrete classes below are *flatten sources*: constructors with
the reference's names and arguments that emit leaves in the reference's child order (child order fixes the
tie-break and the leaf index). Geometry formulas are cited per class.
"""
from __future__ import annotations

from typing import Callable, Union

import numpy as np

from .elements import (BaseRefraciveSurface, CircleRefractive, Lens, OpticalComponent, SphereRefractive,
                       SquareMirror, SquareRefractive, _UNSET)
from .pose import pivot
from .shapes import ASphere, Plane, Polygon, Surface, exact_spherical_asphere, parametric_asphere

_Z = (0, 0, 1)
_Y = (0, 1, 0)


def _pair(value):
    return value if isinstance(value, tuple) else (value, value)


def _grid(ny, nz):
    """(row i along z, column j along y) in the reference's nesting order: z outer, y inner."""
    for i in range(nz):
        for j in range(ny):
            yield i, j


class ComponentGroup(OpticalComponent):
    def __init__(self, origin, **kwargs):
        super().__init__(origin, **kwargs)
        self._bboxes = []
        self.components = []
        self.monitors = []
        self.rays = []
        self.refpoints = []

    def __repr__(self):
        return f"{type(self).__name__}(origin={self.origin}, transform_matrix={self.transform_matrix})"

    # cached boxes (never invalidated, like the reference)
    @property
    def bbox(self):
        if self._bbox == _UNSET:
            self._bbox = tuple(self.get_bbox())
        return tuple(self._bbox)

    @property
    def bboxes(self):
        if not self._bboxes:
            self.get_bboxes()
        return self._bboxes

    def get_bboxes(self):
        self._bboxes = [c.bbox for c in self.components]
        return self._bboxes

    def get_bbox(self) -> tuple:
        return Surface.merge_bboxs(self.get_bboxes())

    # rigid motion of the whole assembly
    def _members(self):
        return (*self.rays, *self.components, *self.monitors, *self.refpoints)

    def _RotAroundLocal(self, axis, localpoint, theta):
        centre = self.origin + np.array(localpoint)
        R = self.R(axis, theta)
        self.transform_matrix = R @ self.transform_matrix
        self.origin = pivot(self.origin, R, localpoint)
        for member in self._members():
            member._RotAroundLocal(axis, centre - member.origin, theta)
        return self

    def _Translate(self, movement):
        self.origin += np.array(movement)
        for member in self._members():
            member._Translate(movement)
        return self

    # membership
    def add_rays(self, rays):
        self.rays.extend(rays)

    def add_component(self, component):
        self.components.append(component)
        if hasattr(component, "rays"):
            self.rays.extend(component.rays)

    def add_components(self, components):
        for component in components:
            self.add_component(component)

    def add_monitor(self, monitor):
        self.monitors.append(monitor)

    def add_monitors(self, monitors):
        self.monitors.extend(monitors)

    def add_refpoint(self, point):
        self.refpoints.append(point)


# ---- slabs and wedges -----------------------------------------------------------------------------------
class GlassSlab(ComponentGroup):
    """Front face n1->n2 at the origin, back face n2->n1 at x = -thickness (component_group.py:148-185)."""

    def __init__(self, origin, width=1.0, height=1.0, thickness=1.0, n1=1.0, n2=1.5, reflectivity=0,
                 transmission=1, **kwargs):
        super().__init__(origin, **kwargs)
        for dx, (na, nb) in ((0, (n1, n2)), (-thickness, (n2, n1))):
            self.add_component(SquareRefractive(origin + np.array([dx, 0, 0]), width, height, na, nb,
                                                reflectivity=reflectivity, transmission=transmission, **kwargs))


class CircleGlassSlab(ComponentGroup):
    """Circular slab with per-face coatings (component_group.py:188-225)."""

    def __init__(self, origin, radius=1.0, thickness=1.0, n1=1.0, n2=1.5, reflectivity1=0, transmission1=1,
                 reflectivity2=0, transmission2=1, **kwargs):
        super().__init__(origin, **kwargs)
        self.radius = radius
        faces = ((0, n1, n2, reflectivity1, transmission1), (-thickness, n2, n1, reflectivity2, transmission2))
        for dx, na, nb, refl, trans in faces:
            self.add_component(CircleRefractive(origin + np.array([dx, 0, 0]), radius, na, nb,
                                                reflectivity=refl, transmission=trans, **kwargs))


class WedgePlate(ComponentGroup):
    """Two faces at +-thickness/2 tilted by +-wedge_angle/2 about z (component_group.py:394-432)."""

    def __init__(self, origin, width=1.0, height=1.0, thickness=1.0, wedge_angle=0.0, n1=1.0, n2=1.5,
                 reflectivity=0, transmission=1, **kwargs):
        super().__init__(origin, **kwargs)
        for sign, (na, nb) in ((+1, (n1, n2)), (-1, (n2, n1))):
            face = SquareRefractive(origin + np.array([sign * thickness / 2, 0, 0]), width, height, na, nb,
                                    reflectivity=reflectivity, transmission=transmission, **kwargs)
            self.add_component(face.RotZ(sign * wedge_angle / 2))


# ---- arrays ---------------------------------------------------------------------------------------------
class MLA(ComponentGroup):
    """ny x nz thin lenses on a square pitch (component_group.py:228-246)."""

    def __init__(self, origin, N, pitch, focal_length, radius, focal_drift=0, **kwargs):
        super().__init__(origin)
        self.pitch, self.focal_length, self.radius = pitch, focal_length, radius
        ny, nz = (N, 1) if isinstance(N, int) else N
        for i, j in _grid(ny, nz):
            centre = np.array([0, (j - (ny - 1) / 2) * pitch, (i - (nz - 1) / 2) * pitch]) + self.origin
            f = focal_length * (1 + focal_drift * np.random.randn())
            self.add_component(Lens(origin=centre, focal_length=f, radius=radius, **kwargs))


def _micro_mirror(vertex, roc, pitch, n, kwargs):
    """Concave micro-mirror: spherical cap of radius `roc` whose pole sits at `vertex`; cap height from the pitch."""
    return SphereRefractive(origin=vertex + [-roc, 0, 0], radius=roc,
                            height=roc - np.sqrt(roc ** 2 - (pitch / 2) ** 2), n1=n, n2=1.0, **kwargs)


class MMA(ComponentGroup):
    """Micro-mirror array: ny x nz spherical caps in a substrate of index n plus its flat back face
    (component_group.py:249-304)."""

    def __init__(self, origin, N, pitch, roc, n, thickness, roc_drift=0, **kwargs):
        super().__init__(origin, **kwargs)
        self.pitch = pitch
        ny, nz = _pair(N) if isinstance(N, tuple) else (N, 1)
        py, pz = _pair(pitch)
        if isinstance(roc, tuple):
            roc = np.array(roc)
            assert roc.shape == (nz, ny), f"roc shape {roc.shape} does not match ({nz}, {ny})"
        else:
            roc = np.ones((nz, ny)) * roc
        shift_y, shift_z = kwargs.get("mma_shifty", 0), kwargs.get("mma_shiftz", 0)
        skew_y, skew_z = kwargs.get("shifty_z", 0), kwargs.get("shiftz_y", 0)
        for i, j in _grid(ny, nz):
            z = (i - (nz - 1) / 2) * pz + (j * skew_z) + shift_z
            y = (j - (ny - 1) / 2) * py + (i * skew_y) + shift_y
            r = roc[i, j] * (1 + roc_drift * np.random.randn())
            self.add_component(_micro_mirror(np.array([0, y, z]) + self.origin, r, self.pitch, n, kwargs))
        self.add_component(SquareRefractive(
            origin=self.origin + np.array([thickness, 0, 0]),
            width=kwargs.get("mma_width", ny * py), height=kwargs.get("mma_height", nz * pz), n1=1, n2=n,
            reflectivity=kwargs.get("back_reflectivity", 0), transmission=kwargs.get("back_transmission", 1)))


def _rotation_between(a, b):
    """(axis, angle) rotating unit vector a onto b (used to aim disordered micro-mirrors). Like the reference
    (solver.py:112-132) a (anti)parallel pair yields the null rotation."""
    a, b = np.asarray(a, float) / np.linalg.norm(a), np.asarray(b, float) / np.linalg.norm(b)
    axis = np.cross(a, b)
    s = np.linalg.norm(axis)
    if s < 1e-12:
        return np.array([1, 0, 0]), 0.0
    return axis / s, float(np.arccos(np.clip(a @ b, -1.0, 1.0)))


class MMADisordered(ComponentGroup):
    """Micro-mirrors at arbitrary positions PList, optionally aimed along -nList[i]
    (component_group.py:307-364)."""

    def __init__(self, origin, PList, pitch, roc, n, thickness, roc_drift=0, nList=None, **kwargs):
        super().__init__(origin, **kwargs)
        self.pitch = pitch
        PList = np.array(PList)
        assert PList.shape[1] == 3, "PList must be a list of 3D points"
        if nList is not None:
            nList = np.array(nList)
            assert nList.shape[0] == PList.shape[0], "nList must match PList length"
        roc = np.ones(PList.shape[0]) * roc if isinstance(roc, (int, float)) else np.array(roc)
        assert roc.shape[0] == PList.shape[0], "roc must match PList length"
        for k, P in enumerate(PList):
            r = roc[k] * (1 + roc_drift * np.random.randn())
            mirror = _micro_mirror(np.array(P) + self.origin, r, self.pitch, n, kwargs)
            if nList is not None:
                axis, theta = _rotation_between(mirror.normal, -nList[k])
                mirror._RotAroundLocal(axis, [r, 0, 0], theta)
            self.add_component(mirror)
        span = PList.max(axis=0) - PList.min(axis=0)
        self.add_component(SquareRefractive(origin=self.origin + np.array([thickness, 0, 0]), width=span[1] * 1.1,
                                            height=span[2] * 1.1, n1=1, n2=n, reflectivity=0, transmission=1))


class DMD(ComponentGroup):
    """ny x nz square mirrors tilted about z (component_group.py:367-391)."""

    def __init__(self, origin, N, pitch, tilt_angle=np.pi / 4, **kwargs):
        super().__init__(origin)
        self.pitch, self.tilt_angle = pitch, tilt_angle
        ny, nz = (N, 1) if isinstance(N, int) else N
        for i, j in _grid(ny, nz):
            centre = np.array([0, (j - (ny - 1) / 2) * pitch, (i - (nz - 1) / 2) * pitch]) + self.origin
            self.add_component(SquareMirror(origin=centre, width=pitch, height=pitch, reflectivity=1.0,
                                            **kwargs).RotZ(self.tilt_angle))


# ---- roof / prism style assemblies: two faces hinged on the apex line through the origin ---------------------
def _hinged_faces(make_face, origin, width, angle):
    """Two faces of length `width` meeting at the origin with included angle `angle`, symmetric about x."""
    tilt = (np.pi - angle) / 2
    upper = make_face(1, origin + np.array([0, width / 2, 0]))._RotAroundLocal(_Z, [0, -width / 2, 0], tilt)
    lower = make_face(2, origin + np.array([0, -width / 2, 0]))._RotAroundLocal(_Z, [0, width / 2, 0], -tilt)
    return upper, lower


class MirrorPair(ComponentGroup):
    """Two square mirrors forming a roof of included angle `angle` (component_group.py:435-497)."""

    def __init__(self, origin, width=1.0, height=1.0, angle: float = np.pi / 2, reflectivity_1=1, transmission_1=0,
                 reflectivity_2=1, transmission_2=0, **kwargs):
        super().__init__(origin, **kwargs)
        self.name = kwargs.get("name", type(self).__name__)
        coat = {1: (reflectivity_1, transmission_1), 2: (reflectivity_2, transmission_2)}
        group_name = self.name

        def face(k, centre):
            kw = dict(kwargs)
            kw["name"] = f"{group_name} mirror {k}"
            return SquareMirror(centre, width=width, height=height, reflectivity=coat[k][0],
                                transmission=coat[k][1], **kw)

        self.add_components(_hinged_faces(face, origin, width, angle))


class MirrorPrism(ComponentGroup):
    """Roof of two identical mirrors (component_group.py:715-747)."""

    def __init__(self, origin, width=1.0, height=1.0, angle: float = np.pi / 2, reflectivity=1.0,
                 transmission=0.0, **kwargs):
        super().__init__(origin, **kwargs)

        def face(k, centre):
            return SquareMirror(centre, width, height, reflectivity=reflectivity, transmission=transmission, **kwargs)

        self.add_components(_hinged_faces(face, origin, width, angle))


class Prism(ComponentGroup):
    """Isosceles prism: two legs of length `width` meeting at the origin plus the base
    (component_group.py:500-597)."""

    def __init__(self, origin, width=1.0, height=1.0, n1=1.0, n2=1.5, angle: float = np.pi / 2, reflectivity_leg=0,
                 transmission_leg=1, reflectivity_hyp=0, transmission_hyp=1, **kwargs):
        super().__init__(origin, **kwargs)
        self.name = kwargs.get("name", type(self).__name__)
        group_name = self.name

        def named(label):
            kw = dict(kwargs)
            kw["name"] = f"{group_name} {label}"
            return kw

        def leg(k, centre):
            return SquareRefractive(centre, width, height, n1, n2, reflectivity=reflectivity_leg,
                                    transmission=transmission_leg, **named(f"leg {k}"))

        self.add_components(_hinged_faces(leg, origin, width, angle))
        self.add_component(SquareRefractive(origin + np.array([-width * np.cos(angle / 2), 0, 0]),
                                            width * 2 * np.sin(angle / 2), height, n2, n1,
                                            reflectivity=reflectivity_hyp, transmission=transmission_hyp,
                                            **named("hypotenuse")))


class TriangularPrism(ComponentGroup):
    """Entrance face 1 of length `width` normal to x with the origin at its lower end; face 2 leaves its upper
    end at angle alpha, face 3 its lower end at angle beta. Faces 2 and 3 carry interact caps
    (component_group.py:600-712)."""

    def __init__(self, origin, width=1.0, height=1.0, n1=1.0, n2=1.5, alpha=np.pi / 4, beta=np.pi / 2,
                 reflectivity_1=0, reflectivity_2=0, reflectivity_3=0, transmission_1=1, transmission_2=1,
                 transmission_3=1, max_interact_count_2=5, max_interact_count_3=5, **kwargs):
        super().__init__(origin, **kwargs)
        self.name = kwargs.get("name", type(self).__name__)
        apex = np.sin(np.pi - alpha - beta)
        len2, len3 = width * np.sin(beta) / apex, width * np.sin(alpha) / apex  # law of sines
        self.add_component(SquareRefractive(origin=origin + np.array([0, width / 2, 0]), width=width, height=height,
                                            n1=n1, n2=n2, reflectivity=reflectivity_1, transmission=transmission_1,
                                            **kwargs))
        self.add_component(SquareRefractive(origin=origin + np.array([0, width - len2 / 2, 0]), width=len2,
                                            height=height, n1=n2, n2=n1, reflectivity=reflectivity_2,
                                            transmission=transmission_2, max_interact_count=max_interact_count_2,
                                            **kwargs)._RotAroundLocal(_Z, [0, len2 / 2, 0], -alpha))
        self.add_component(SquareRefractive(origin=origin + np.array([0, len3 / 2, 0]), width=len3, height=height,
                                            n1=n2, n2=n1, reflectivity=reflectivity_3, transmission=transmission_3,
                                            max_interact_count=max_interact_count_3,
                                            **kwargs)._RotAroundLocal(_Z, [0, -len3 / 2, 0], beta))


class MirrorCube(ComponentGroup):
    """Corner cube of three L x L mirrors, turned so the x axis makes equal angles with all three
    (component_group.py:750-783)."""

    def __init__(self, origin, L=1.0, reflectivity=1.0, **kwargs):
        super().__init__(origin, **kwargs)
        h = L / 2

        def wall(offset):
            return SquareMirror(self.origin + np.array(offset), width=L, height=L, reflectivity=reflectivity, **kwargs)

        self.add_components([wall([0.0, h, h]), wall([h, 0.0, h]).RotZ(np.pi / 2), wall([h, h, 0.0]).RotY(-np.pi / 2)])
        self._RotAroundLocal(_Z, [0, 0, 0], -np.pi / 4)
        self._RotAroundLocal(_Y, [0, 0, 0], np.arccos(np.sqrt(2 / 3)))


class DovePrism(ComponentGroup):
    """Dove prism of base length L, height D, index Ng: two trapezoidal side faces (2-D polygons), top and
    bottom rectangles, and the two slanted end facets (3-D polygons) (component_group.py:786-850).
    Children are placed in absolute coordinates exactly as the reference does (the `origin` argument only
    seeds the group pose)."""

    def __init__(self, origin, L, D, Ng, **kwargs):
        super().__init__(origin, **kwargs)
        self.L, self.D, self.Ng = L, D, Ng
        top = L - 2 * D
        trapezoid = Polygon(np.array([[-L / 2, 0], [L / 2, 0], [top / 2, D], [-top / 2, D]]))

        def facet(surface, n1, n2, at=(0, 0, 0)):
            face = BaseRefraciveSurface(origin=list(at), n1=n1, n2=n2)
            face.surface = surface
            return face

        def end(sign):
            y0, y1 = sign * L / 2, sign * top / 2
            return Polygon(np.array([[-D / 2, y0, 0], [D / 2, y0, 0], [D / 2, y1, D], [-D / 2, y1, D]]))

        side_l = facet(trapezoid, Ng, 1, at=(-D / 2, 0, 0))
        side_r = facet(trapezoid, 1, Ng, at=(D / 2, 0, 0))
        roof = SquareRefractive(origin=[0, 0, D], width=top, height=D, n1=Ng, n2=1).RotY(np.pi / 2)
        base = SquareRefractive(origin=[0, 0, 0], width=L, height=D, n1=1, n2=Ng).RotY(np.pi / 2)
        self.add_components([side_l, side_r, roof, base, facet(end(-1), Ng, 1), facet(end(+1), 1, Ng)])

    @property
    def z0(self):
        """Height at which an axial ray leaves undeviated."""
        theta = np.arcsin((np.sqrt(2) / 2) / self.Ng)
        return (self.L / 2) / (1 + np.tan(np.pi / 4 + theta))


# ---- thick lenses ----------------------------------------------------------------------------------------
def _sag(R, diameter):
    return abs(R) - np.sqrt(R ** 2 - (diameter / 2) ** 2)


def _spherical_face(vertex_x, R, diameter, n_before, n_after, origin, kwargs):
    """Spherical interface whose pole is at x = vertex_x (relative to `origin`), radius R signed as seen by a
    beam travelling +x (R > 0: centre of curvature behind the pole). The cap of a SphereRefractive faces its
    local +x, so the R > 0 case is the cap turned by pi about z; n1 is always the medium on the cap's +x side."""
    centre = np.array([R + vertex_x, 0, 0]) + origin
    if R > 0:
        return SphereRefractive(origin=centre, radius=R, height=_sag(R, diameter), n1=n_before, n2=n_after,
                                **kwargs).RotZ(np.pi)
    return SphereRefractive(origin=centre, radius=-R, height=_sag(R, diameter), n1=n_after, n2=n_before, **kwargs)


class PlanoConvexLens(ComponentGroup):
    """Curved face towards -x, flat face towards +x; index derived from n = 1 + R/EFL; the origin is the
    principal plane, CT/n in front of the flat face (component_group.py:853-878)."""

    def __init__(self, origin, EFL, CT, diameter, R, **kwargs):
        super().__init__(origin, **kwargs)
        n = 1 + R / EFL
        flat_x = CT / n
        self.add_component(SphereRefractive(origin=np.array([R - (CT - flat_x), 0, 0]) + self.origin, radius=R,
                                            height=_sag(R, diameter), n1=1.0, n2=n, **kwargs).RotZ(np.pi))
        self.add_component(CircleRefractive(origin=self.origin + np.array([+flat_x, 0, 0]), radius=diameter / 2,
                                            n1=n, n2=1.0, **kwargs).RotZ(np.pi))


class BiConvexLens(ComponentGroup):
    """Two spherical faces, R1 at the origin and R2 at x = CT (signed for a +x beam). Without `n` the index is
    solved from the lensmaker equation for the given EFL (component_group.py:881-938)."""

    def __init__(self, origin, CT, R1, R2, diameter, EFL=None, n=None, **kwargs):
        super().__init__(origin, **kwargs)
        if n is None:
            assert EFL is not None, "Either n or EFL must be provided"
            n = 1 + (1 / EFL) / (1 / R1 - 1 / R2)
            for _ in range(3):  # thick-lens correction, fixed-point
                n = 1 + (1 / EFL) / (1 / R1 - 1 / R2 + (n - 1) * CT / (n * R1 * R2))
        else:
            assert isinstance(n, (int, float)), "n must be a number"
        self.add_component(_spherical_face(0, R1, diameter, 1.0, n, self.origin, kwargs))
        self.add_component(_spherical_face(CT, R2, diameter, n, 1.0, self.origin, kwargs))


class Doublet(ComponentGroup):
    """Cemented doublet: faces R1 | glass n12 | R2 | glass n23 | R3 (component_group.py:941-1011)."""

    def __init__(self, origin, CT1, CT2, R1, R2, R3, diameter, n12, n23, **kwargs):
        super().__init__(origin, **kwargs)
        self.add_component(_spherical_face(0, R1, diameter, 1.0, n12, self.origin, kwargs))
        self.add_component(_spherical_face(CT1, R2, diameter, n12, n23, self.origin, kwargs))
        self.add_component(_spherical_face(CT1 + CT2, R3, diameter, n23, 1.0, self.origin, kwargs))


class ASphericLens(ComponentGroup):
    """Aspheric front face x = f_asphere_1(r) at the origin and, at x = CT, either a second asphere or a flat
    face (component_group.py:1014-1062). Both faces are turned by pi about z (normals towards -x)."""

    def __init__(self, origin, CT, f_asphere_1: Callable, f_asphere_2: Union[Callable, None], diameter, n, **kwargs):
        super().__init__(origin, **kwargs)
        self.f_asphere_1, self.f_asphere_2 = f_asphere_1, f_asphere_2
        front = BaseRefraciveSurface(origin=self.origin, n1=1, n2=n, surface=ASphere(diameter / 2, f_asphere_1),
                                     **kwargs).RotZ(np.pi)
        back_at = self.origin + np.array([CT, 0, 0])
        if f_asphere_2 is not None:
            back = BaseRefraciveSurface(origin=back_at, n1=n, n2=1.0, surface=ASphere(diameter / 2, f_asphere_2), **kwargs)
        else:
            back = CircleRefractive(origin=back_at, radius=diameter / 2, n1=n, n2=1.0, **kwargs)
        self.add_components([front, back.RotZ(np.pi)])


class ASphericExactSphericalLens(ASphericLens):
    def __init__(self, origin, EFL, CT, diameter, n, **kwargs):
        super().__init__(origin, CT, exact_spherical_asphere(EFL, n), None, diameter, n, **kwargs)


class ASphericParametricLens(ASphericLens):
    def __init__(self, origin, CT, diameter, n, R, kappa, a4=0, a6=0, a8=0, **kwargs):
        super().__init__(origin, CT, parametric_asphere(R, kappa, a4, a6, a8), None, diameter, n, **kwargs)
