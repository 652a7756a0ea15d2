"""Scene fixtures shared by the parity tests, the golden-vector generator and the benchmarks.

Every builder takes a namespace `ns` exposing the optable public names (Ray, Mirror, Lens, ...): the
reference package in the build container, optable_b200 everywhere. The geometry restates the
reference's example scripts (examples/*.py, cited per builder) and SURVEY.md section 8(d).
"""
from __future__ import annotations

import numpy as np

from optable_b200.workloads import SEED, Scene, asphere_lens9, cavity, ripa, telescope_4f  # noqa: F401  (shared with bench.py)


def gaussian_beam(ns):
    """C1: examples/gaussian_beam.py:16-44 (README scene)."""
    wl, w0 = 780e-9, 10e-6
    rays = [ns.Ray([-10, y, 0], [1, 0, 0], wavelength=wl, w0=w0) for y in (0, 2, 4, 6, 9)]
    rays.append(ns.Ray([-10, 21, 0], [1, 0, 0], wavelength=wl, w0=w0).RotZ(-np.pi / 4))
    comps = [
        ns.Mirror([0, 0, 0]).RotZ(np.pi / 6),
        ns.Lens([0, 2, 0], radius=0.8, focal_length=5),
        ns.Lens([0, 4, 0], radius=0.8, focal_length=10),
        ns.Lens([0, 6.5, 0], radius=0.8, focal_length=10),
        ns.GlassSlab([0, 9, 0], n1=1, n2=2, thickness=5),
        ns.Mirror([0, 11, 0]).RotZ(-np.pi / 2),
    ]
    return Scene(comps, rays)


def glass_slab(ns):
    """examples/glass_slab.py:34-52: splitting at both faces; the 2000-pop cap binds."""
    rays = [ns.Ray([-3, 2, 0], [np.cos(np.pi / 6), -np.sin(np.pi / 6), 0], wavelength=780e-7, w0=2e-4).Propagate(-2)]
    gs = ns.GlassSlab([0, 0, 0], width=2, height=2, thickness=0.5, n1=1, n2=1.5, reflectivity=0.2)
    return Scene([gs], rays, limit={"max_trace_num": 300})


def chromatic(ns):
    """examples/chromatic_aberration.py:16-40: 3 wavelengths share one _id; Sellmeier glass."""
    r0 = [ns.Ray([-3, 2, 0], [np.cos(np.pi / 6), -np.sin(np.pi / 6), 0], wavelength=780e-7, w0=20e-4).Propagate(-2)]
    rays = ns.multiplex_rays_in_wavelength(r0, [780e-7, 560e-7, 400e-7])
    gs = ns.GlassSlab([0, 0, 0], width=2, height=2, thickness=0.5, n1=ns.Vacuum(), n2=ns.Glass_NBK7(), reflectivity=0.2)
    return Scene([gs], rays, limit={"max_trace_num": 200})


def cavity_aligned(ns):
    return cavity(ns, 0.0, 0.0, gaussian=True, n_rays=3, limit={"max_trace_num": 400})


def doublet(ns, n_rays=6):
    """Edmund #88-597 doublet (examples/calibrate_4f.py:126-147) + monitor, 3 wavelengths."""
    lens = ns.Doublet([30.3964, 0, 0], CT1=1.359, CT2=0.6, R1=18.405, R2=-13.734, R3=-39.933,
                      n12=ns.Glass_NBK7(), n23=ns.Glass_NSF5(), diameter=7.5)
    mon = ns.Monitor([62, 0, 0], 10, 10)
    rng = np.random.default_rng(SEED + 1)
    r0 = [ns.Ray([0, 2, -1], [1, 0, 0], wavelength=780e-7, w0=61e-4)]
    for _ in range(n_rays - 1):
        y, z = rng.uniform(-3, 3, 2)
        r0.append(ns.Ray([0, y, z], [1, 0.01 * rng.standard_normal(), 0.01 * rng.standard_normal()],
                         wavelength=780e-7, w0=61e-4))
    rays = ns.multiplex_rays_in_wavelength(r0, [780e-7, 560e-7, 400e-7])
    return Scene([lens], rays, [mon])


def abcd_rays(ns, N=3, D=0.6):
    """examples/calibrate_4f.py:184-191: 2N+1 Gaussian rays with explicit integer ids."""
    return [ns.Ray([-10, i * D, 0], [1, 0, 0], wavelength=780e-7, w0=61e-4, id=int(i + N)).Propagate(-10)
            for i in np.arange(-N, N + 1)]


def exact_asphere(ns):
    """ASphericExactSphericalLens (component_group.py:1065-1082) with tilted rays."""
    lens = ns.ASphericExactSphericalLens([10, 0, 0], EFL=20.0, CT=0.6, diameter=5.0, n=1.5)
    mon = ns.Monitor([30, 0, 0], 6, 6)
    rng = np.random.default_rng(SEED + 3)
    rays = [ns.Ray([0, *rng.uniform(-2, 2, 2)], [1, *(0.02 * rng.standard_normal(2))], wavelength=633e-7, w0=50e-4)
            for _ in range(8)]
    return Scene([lens], rays, [mon])


def mirror_pair(ns):
    """examples/mirror_pair.py:20-36."""
    r0 = ns.Ray([-10, 1, 0], [1, 0, 0])._RotAround([0, 1, 0], [0, 0, 0], 0.1)
    mp = ns.MirrorPair([3, 0, 0], 4, 4).RotX(0.3)
    return Scene([mp], [r0], [ns.Monitor([1, 0, 0], 5, 5), ns.Monitor([-3, 0, 0], 5, 5)])


def prism_refl(ns):
    """examples/prism_refl.py:21-65: TriangularPrism with interact caps and explicit ray ids."""
    theta, L, wl, w0 = 0.00956, 6, 780e-7, 61e-4
    n = ns.Glass_NBK7().n(780e-9)
    D = 3 - 210e-4
    rays = [ns.Ray([3, y + L / 2, 0], [-1, 0, 0], wavelength=wl, w0=w0, id=i).Propagate(-3).RotZ(theta)
            for i, y in enumerate(np.linspace(-D, D, 5))]
    ps = ns.TriangularPrism(origin=[0, 0, 0], width=L, height=L, n1=1, n2=n, alpha=np.pi / 4, beta=np.pi / 2,
                            reflectivity_1=1, reflectivity_3=0.01, max_interact_count_2=10, max_interact_count_3=10)
    return Scene([ps], rays, [ns.Monitor([-3, 0, 0], width=L, height=L)], limit={"max_trace_num": 300})


def dove_prism(ns):
    """examples/dove_prism.py:31-66: 2-D and 3-D polygon faces, TIR on the base."""
    L, D, Ng = 6.34, 1.515, 1.515
    dp = ns.DovePrism([0, 0, 0], L=L, D=D, Ng=Ng)
    z0 = dp.z0
    rays = [ns.Ray([x, -50, z0], [0, 1, 0]) for x in np.linspace(-0.6, 0.6, 7)]
    rays += [ns.Ray([x, 3, z], [0, -1, 0]) for x in np.linspace(-1, 1, 3) for z in np.linspace(-1, 1, 3)]
    rays += [ns.Ray([3, y, z + 1], [-1, 0, -0.3]) for y in np.linspace(-1, 1, 3) for z in np.linspace(-1, 1, 3)]
    dp = dp.RotX(0.02).RotZ(-0.01).TX(0.1).TZ(0.05).RotYAroundLocal([0, 0, D / 2], theta=0.3)
    mon0 = ns.Monitor([0, 10, z0], width=2 * D, height=2 * D).RotZ(np.pi / 2)
    return Scene([dp], rays, [mon0])


def gaussian_telescope(ns):
    """examples/gaussian_telescope.py:20-49 (thin lenses; `diameter=` kwarg is ignored by Lens)."""
    F, M, D = 5, 5, 2
    F0, F1, F2, F3 = F, F / M, F * M, F
    l0 = ns.Lens([F0, 0, 0], focal_length=F0, diameter=D)
    l2 = ns.Lens([2 * F0 + 2 * F1 + F2, 0, 0], focal_length=F2, diameter=D)
    l3 = ns.Lens([2 * F0 + 2 * F1 + 2 * F2 + F3, 0, 0], focal_length=F3, diameter=D)
    mon0 = ns.Monitor([2 * F0 + 2 * F1 + 2 * F2 + 2 * F3, 0, 0], width=D, height=D)
    rays = [ns.Ray([0, 0.1 * k, 0.05 * k], [1, 0, 0], wavelength=780e-7, w0=61e-4) for k in range(4)]
    return Scene([l0, l2, l3], rays, [mon0])


def misc_components(ns):
    """Beam splitter, block with a circular hole, cylindrical mirror, wedge, plano-convex, bi-convex, MLA, DMD,
    MirrorCube, Prism, CircleGlassSlab, finite-length and dead input rays."""
    comps = [
        ns.BeamSplitter([2, 0, 0], width=2, height=2, eta=0.3).RotZ(np.pi / 4),
        ns.Block([6, 0, 0], hole=ns.Circle(0.5), width=3, height=3),
        ns.CylMirror([12, 0.2, 0], radius=2.0, height=3.0, theta_range=(-np.pi / 3, np.pi / 3)),
        ns.WedgePlate([2, 5, 0], width=2, height=2, thickness=0.4, wedge_angle=0.05, n1=1.0, n2=1.45),
        ns.PlanoConvexLens([2, 9, 0], EFL=10.0, CT=0.5, diameter=2.0, R=5.0),
        ns.BiConvexLens([8, 9, 0], CT=0.6, R1=12.0, R2=-12.0, diameter=2.0, n=1.5),
        ns.MLA([14, 9, 0], N=(3, 2), pitch=0.5, focal_length=4.0, radius=0.25),
        ns.DMD([2, -5, 0], N=(2, 2), pitch=0.6, tilt_angle=np.pi / 4 + 0.1),
        ns.MirrorCube([8, -6, 0], L=2.0),
        ns.Prism([5, -12, 0], width=2, height=2, n1=1.0, n2=1.5, reflectivity_hyp=0.1),
        ns.CircleGlassSlab([2, 14, 0], radius=1.0, thickness=0.3, n1=1.0, n2=1.6, reflectivity2=0.05),
        ns.SquareMirror([20, 0, 0], width=30, height=30, reflectivity=0.5, transmission=0.5),
    ]
    rng = np.random.default_rng(SEED + 4)
    rays = []
    for y0 in (0.0, 5.0, 9.0, -5.0, -6.0, -12.0, 14.0):
        for _ in range(4):
            dy, dz = rng.uniform(-0.3, 0.3, 2)
            rays.append(ns.Ray([-3, y0 + dy, dz], [1, 0.03 * rng.standard_normal(), 0.03 * rng.standard_normal()],
                               wavelength=633e-7, w0=30e-4))
    rays.append(ns.Ray([-3, 0.1, 0], [1, 0, 0], length=4.0))           # limit before the block
    rays.append(ns.Ray([-3, 0.2, 0], [1, 0, 0], length=7.0, wavelength=500e-7))
    rays.append(ns.Ray([-3, 0.3, 0], [1, 0, 0], alive=False))
    rays.append(ns.Ray([-3, 0.9, 0.0], [1, 0, 0]))                      # no q, no wavelength
    mons = [ns.Monitor([18, 0, 0], 40, 40), ns.Monitor([-2, 0, 0], 40, 40).RotZ(0.2)]
    return Scene(comps, rays, mons, limit={"max_trace_num": 60})


def mma_small(ns):
    """A small MMA cavity in the style of examples/ripa_gen2_lensless.py:176-200 (nested groups)."""
    pitch, roc = 420e-4, 2.2
    m0 = ns.MMA(origin=[-0.1, 0, 0], N=(6, 1), pitch=pitch, roc=roc, n=1.5, thickness=0.1, reflectivity=0.9999,
                transmission=0, back_transmission=0, back_reflectivity=1)
    m1 = ns.MMA(origin=[2.0, 0, 0], N=(6, 3), pitch=pitch, roc=roc, n=1.5, thickness=0.1, reflectivity=0.98,
                transmission=0.3, shifty_z=-0.002).TY(0.01)
    grp = ns.ComponentGroup([0, 0, 0])
    grp.add_components([m0, m1])
    mon = ns.Monitor([2.0 - 1e-4, 0, 0], width=1, height=1)
    rng = np.random.default_rng(SEED + 5)
    rays = []
    for _ in range(6):
        y, z = rng.uniform(-0.1, 0.1, 2)
        rays.append(ns.Ray([0, y, z], [1, 0.02 * rng.standard_normal(), 0.02 * rng.standard_normal()],
                           wavelength=780e-7, w0=40e-4))
    return Scene([grp], rays, [mon], limit={"max_trace_num": 120})


def extras(ns):
    """Corner cases the example scenes do not reach: a union aperture (Plane.union, surfaces.py:100-117), a
    non-orthonormal transform_matrix (sheared/scaled by hand: the to-local normalisation really matters), a diverging
    thin lens, a full-circle cylinder mirror hit from inside, a polygonal mirror, rays without wavelength and with a
    length limit that ends between two surfaces."""
    m_union = ns.SquareMirror([4, 0, 0], width=1, height=1, reflectivity=0.7, transmission=0.3)
    m_union.surface = ns.Circle(0.6).union(ns.Rectangle(2.4, 0.3))
    skew = ns.SquareRefractive([7, 0.2, 0], width=3, height=3, n1=1.0, n2=1.4, reflectivity=0.1).RotZ(0.2)
    skew.transform_matrix = skew.transform_matrix @ np.array([[1.0, 0.15, 0.0], [0.0, 1.3, 0.1], [0.0, 0.0, 0.8]])
    lens = ns.Lens([10, 0, 0], focal_length=-3.0, radius=1.5, transmission=0.9)
    cyl = ns.CylMirror([16, 0, 0], radius=2.5, height=4.0)
    tri = ns.BaseMirror([-4, 0, 0], reflectivity=1.0)
    tri.surface = ns.Polygon([[-1.0, -1.2], [1.4, -0.8], [0.2, 1.5]])
    tri.RotZ(np.pi).RotY(0.15)
    rng = np.random.default_rng(SEED + 7)
    rays = []
    for k in range(14):
        y, z = rng.uniform(-0.5, 0.5, 2)
        kw = dict(wavelength=633e-7, w0=40e-4) if k % 3 else {}
        rays.append(ns.Ray([0, y, z], [1, 0.05 * rng.standard_normal(), 0.05 * rng.standard_normal()], **kw))
    rays.append(ns.Ray([0, 0.1, 0.1], [1, 0, 0], length=5.5, wavelength=500e-7, w0=20e-4))
    rays.append(ns.Ray([0, -0.1, 0.2], [-1, 0.02, 0.01]))
    return Scene([m_union, skew, lens, cyl, tri], rays, [ns.Monitor([12, 0, 0], 8, 8).RotY(0.1)], limit={"max_trace_num": 40})


def nested_csg(ns):
    """Composite apertures whose operands are composites themselves (Plane.union / Plane.subtract compose freely,
    surfaces.py:100-136): a plate with a keyhole (rectangle minus (circle or slot)), an annulus with a bar ((circle minus
    circle) or rectangle), and three levels with a polygon; one of them inside a ComponentGroup (merged boxes)."""
    keyhole = ns.SquareMirror([3, 0, 0], width=1, height=1, reflectivity=0.6, transmission=0.4)
    keyhole.surface = ns.Rectangle(3.0, 3.0).subtract(ns.Circle(0.5).union(ns.Rectangle(0.3, 2.0)))
    annulus = ns.SquareRefractive([6, 0, 0], width=1, height=1, n1=1.0, n2=1.5, reflectivity=0.05)
    annulus.surface = ns.Circle(1.2).subtract(ns.Circle(0.4)).union(ns.Rectangle(3.0, 0.25))
    deep = ns.SquareMirror([0, 0, 0], width=1, height=1, reflectivity=1.0)
    tri = ns.Polygon([[-0.6, -0.5], [0.7, -0.4], [0.1, 0.8]])
    deep.surface = ns.Rectangle(2.0, 1.0).union(ns.Circle(0.9)).subtract(tri).subtract(ns.Circle(0.15).union(ns.Rectangle(0.1, 1.6)))
    grp = ns.ComponentGroup([10, 0, 0])
    grp.add_component(deep)
    grp.RotZ(np.pi + 0.05)
    rng = np.random.default_rng(SEED + 11)
    rays = []
    for _ in range(40):
        y, z = rng.uniform(-1.6, 1.6, 2)
        rays.append(ns.Ray([0, y, z], [1, 0.01 * rng.standard_normal(), 0.01 * rng.standard_normal()], wavelength=633e-7, w0=30e-4))
    return Scene([keyhole, annulus, grp], rays, [ns.Monitor([8, 0, 0], 6, 6), ns.Monitor([-1, 0, 0], 6, 6)], limit={"max_trace_num": 40})


def random_csg(ns, seed, depth=4, n_rays=48):
    """One absorbing/reflecting/refracting plate per scene whose aperture is a random composite of circles, rectangles
    and planar polygons, nested up to `depth` levels of Plane.union / Plane.subtract (surfaces.py:100-136), hit by a
    grid-like bundle; a monitor behind it records what gets through."""
    rng = np.random.default_rng(SEED + 5000 + seed)

    def shape(d):
        if d == 0 or rng.random() < 0.25:
            k = rng.integers(3)
            if k == 0:
                return ns.Circle(float(rng.uniform(0.2, 1.4)))
            if k == 1:
                return ns.Rectangle(float(rng.uniform(0.2, 2.6)), float(rng.uniform(0.2, 2.6)))
            c = rng.uniform(-0.6, 0.6, 2)
            ang = np.sort(rng.uniform(0, 2 * np.pi, 3 + int(rng.integers(3))))
            rad = rng.uniform(0.3, 1.2, len(ang))
            return ns.Polygon([[float(c[0] + r * np.cos(a)), float(c[1] + r * np.sin(a))] for r, a in zip(rad, ang)])
        a, b = shape(d - 1), shape(d - 1)
        return a.union(b) if rng.random() < 0.5 else a.subtract(b)

    kind = int(rng.integers(3))
    if kind == 0:
        plate = ns.SquareMirror([4, 0, 0], width=1, height=1, reflectivity=0.5, transmission=0.5)
    elif kind == 1:
        plate = ns.SquareRefractive([4, 0, 0], width=1, height=1, n1=1.0, n2=1.5, reflectivity=0.1)
    else:
        plate = ns.Block([4, 0, 0], width=1, height=1)
    top = shape(depth)
    while type(top).__name__ != "Plane":   # at least one operator at the top
        top = top.union(shape(depth - 1)) if rng.random() < 0.5 else top.subtract(shape(depth - 1))
    plate.surface = top
    plate.RotZ(float(rng.uniform(-0.2, 0.2))).RotY(float(rng.uniform(-0.2, 0.2)))
    rays = []
    for _ in range(n_rays):
        y, z = rng.uniform(-1.5, 1.5, 2)
        rays.append(ns.Ray([0, y, z], [1, 0.02 * rng.standard_normal(), 0.02 * rng.standard_normal()], wavelength=633e-7, w0=30e-4))
    return Scene([plate], rays, [ns.Monitor([7, 0, 0], 8, 8), ns.Monitor([-1, 0, 0], 8, 8)], limit={"max_trace_num": 20})


def caps_binding(ns):
    """Interact caps that really bind: a TriangularPrism whose faces 2/3 stop interacting after 3 hits
    per ray id, partial reflections everywhere (splitting), and each ray multiplexed into 6 wavelengths that share
    its `_id`, so the visible set depends on the reference's sequential order across wavelengths and pops."""
    L = 6
    ps = ns.TriangularPrism(origin=[0, 0, 0], width=L, height=L, n1=1, n2=ns.Glass_NBK7(), alpha=np.pi / 4, beta=np.pi / 2,
                            reflectivity_1=0.3, reflectivity_2=0.5, reflectivity_3=0.4, max_interact_count_2=3,
                            max_interact_count_3=3)
    r0 = [ns.Ray([3, y + L / 2, 0.1 * y], [-1, 0.02 * y, 0], wavelength=780e-7, w0=61e-4, id=10 + i)
          for i, y in enumerate(np.linspace(-2.5, 2.5, 4))]
    rays = ns.multiplex_rays_in_wavelength(r0, list(np.linspace(450e-7, 900e-7, 6)))
    return Scene([ps], rays, [ns.Monitor([-3, 0, 0], width=2 * L, height=2 * L)], limit={"max_trace_num": 80})


def ripa2_simplified(ns, spacing=10):
    """examples/ripa_gen2_2nd_simplified.py:26-131: two tilted fans of Gaussian rays with explicit ids (reverse ray
    tracing) through a 4f pair of parametric aspheres (Sellmeier UVFS), everything held by one ComponentGroup
    (add_rays / add_components / add_monitors, TY translations, Propagate, _RotAroundLocal)."""
    BW = 6.834682
    wl, DMLA, NMLA = 780e-7, 420e-4, 80
    L = DMLA * NMLA / 2
    d = 3e8 / (2 * BW * 1e9) / 0.01
    w0 = np.sqrt(wl * (2 * d) / (2 * np.pi))
    theta0 = np.arctan(DMLA / d / 2)
    kappa, a4, a6 = -1.01705, -4.00e-11, 3.180e-15

    def fan(idxs, sign, offset=0):
        return [ns.Ray([0, L - (i + offset) * DMLA, 0.0], [1, 0, 0], wavelength=wl, w0=w0, id=int(k + NMLA * int(sign == -1)))
                .Propagate(0.0)._RotAroundLocal([0, 0, 1], [0, 0, 0], -sign * theta0) for k, i in enumerate(idxs)]

    rays = fan(range(0, NMLA, spacing), 1) + fan(range(0, NMLA, spacing), -1, offset=1)
    EFL, CT = 43.17, 0.8
    glass = ns.Glass_UVFS()
    R = EFL * (glass.n(780e-9) - 1)
    lens = dict(CT=CT, diameter=2.54 * 3, R=R, n=glass, kappa=kappa, a4=a4 * (1e-2 / 1e-3) ** 4, a6=a6 * (1e-2 / 1e-3) ** 6)
    l0 = ns.ASphericParametricLens([EFL, 0, 0], name="L0", **lens).TY(DMLA / 2)
    l1 = ns.ASphericParametricLens([3 * EFL, 0, 0], name="L1", **lens).RotZ(np.pi).TY(DMLA / 2)
    mons = [ns.Monitor([-0.5, -2.5, 0], width=5, height=5, name="R2 Monitor 0").TY(L),
            ns.Monitor([4 * EFL, 0, 0], width=5, height=5, name="R2 Monitor 1"),
            ns.Monitor([4 * EFL + d, 0, 0], width=5, height=5, name="R2 Monitor 2")]
    group = ns.ComponentGroup([0, 0, 0], name="RIPA2")
    group.add_rays(rays)
    group.add_components([l0, l1])
    group.add_monitors(mons)
    return Scene([group], group.rays, group.monitors)


def ripa2_postprocess(ns, table, scene):
    """The analysis part of examples/ripa_gen2_2nd_simplified.py:169-203 on a traced table: pair every ray of the
    first fan with its partner (id + 80) at the last monitor, intersect the pair, and derive the mirror position,
    normal, radius of curvature and mean optical path length at the crossing. Returns arrays (one row per pair)."""
    mon2 = scene.monitors[2]
    d = 3e8 / (2 * 6.834682 * 1e9) / 0.01
    first = [r for r in mon2.get_rays(sort="ID") if r._id < 80]
    P, nrm, roc, pl = [], [], [], []
    for ray0 in first:
        ray1, _ = mon2.get_ray_id(ray0._id + 80)
        t0, t1, Pk, nk = ns.solve_ray_ray_intersection(ray0.origin, ray0.direction, ray1.origin, ray1.direction)
        d0, d1 = ray0.distance_to_waist(ray0.q_at_z(t0)), ray1.distance_to_waist(ray1.q_at_z(t1))
        P.append(Pk)
        nrm.append(nk)
        roc.append(2 * (d ** 2 + d0 * d1) / (d0 + d1))
        pl.append((ray0.pathlength(t0) + ray1.pathlength(t1)) / 2)
    return {"P": np.array(P), "n": np.array(nrm), "roc": np.array(roc), "pathlength": np.array(pl)}


def fuzz(ns, seed, n_rays=24, caps=False, extended=False):
    """Random scene for the fuzz parity tests: 4-9 components drawn from the whole component zoo with random
    parameters, positions in a 15 x 5 x 1.6 box, mostly facing the beam, 2 monitors, a cone of rays from the
    left (3 wavelengths, some without q, some length-limited). Deterministic in `seed`; both packages build the
    same scene because all randomness is drawn here."""
    rng = np.random.default_rng(SEED * 7 + seed)
    U = rng.uniform

    def place(c):
        # mostly facing the beam (within +-70 degrees), one in five at an arbitrary azimuth
        az = U(-np.pi, np.pi) if U() < 0.2 else U(-1.2, 1.2)
        return c.RotZ(az).RotY(U(-0.4, 0.4)).RotX(U(-0.3, 0.3))

    def pos():
        return [U(3, 18), U(-2.5, 2.5), U(-0.8, 0.8)]

    glass = [lambda: 1.0 + U(0.3, 0.9), lambda: ns.Glass_NBK7(), lambda: ns.Glass_NSF5() if hasattr(ns, "Glass_NSF5") else 1.6]
    palette = [
        lambda: ns.Mirror(pos(), radius=U(0.5, 2.5), reflectivity=U(0.3, 1.0), transmission=U(0.0, 0.5)),
        lambda: ns.SquareMirror(pos(), width=U(1, 4), height=U(1, 4), reflectivity=U(0.5, 1.0)),
        lambda: ns.BeamSplitter(pos(), width=U(1, 4), height=U(1, 4), eta=U(0.2, 0.8)),
        lambda: ns.CylMirror(pos(), radius=U(1.0, 3.0), height=U(2, 4), theta_range=(-U(0.5, 3.1), U(0.5, 3.1))),
        lambda: ns.Lens(pos(), focal_length=U(-8, 8) or 5.0, radius=U(0.8, 2.5), transmission=U(0.7, 1.0)),
        lambda: ns.Block(pos(), hole=ns.Circle(U(0.2, 0.8)), width=U(2, 4), height=U(2, 4)),
        lambda: ns.GlassSlab(pos(), width=U(2, 4), height=U(2, 4), thickness=U(0.2, 1.5), n1=1.0, n2=rng.choice(glass)(),
                             reflectivity=U(0.0, 0.2)),
        lambda: ns.CircleGlassSlab(pos(), radius=U(0.8, 2.0), thickness=U(0.2, 1.0), n1=1.0, n2=1.0 + U(0.3, 0.8),
                                   reflectivity2=U(0.0, 0.1)),
        lambda: ns.WedgePlate(pos(), width=U(2, 4), height=U(2, 4), thickness=U(0.3, 0.8), wedge_angle=U(0.0, 0.1), n1=1.0,
                              n2=1.0 + U(0.3, 0.7)),
        lambda: ns.PlanoConvexLens(pos(), EFL=U(6, 20), CT=U(0.4, 0.8), diameter=U(1.5, 3.0), R=U(4, 10)),
        lambda: ns.BiConvexLens(pos(), CT=U(0.5, 0.9), R1=U(8, 20), R2=-U(8, 20), diameter=U(1.5, 3.0), n=1.0 + U(0.4, 0.8)),
        lambda: ns.Doublet(pos(), CT1=U(0.8, 1.4), CT2=U(0.4, 0.8), R1=U(14, 22), R2=-U(10, 16), R3=-U(30, 50),
                           diameter=U(4, 7), n12=ns.Glass_NBK7(), n23=1.0 + U(0.5, 0.8)),
        lambda: ns.Prism(pos(), width=U(1.5, 3), height=U(1.5, 3), n1=1.0, n2=1.0 + U(0.4, 0.7), reflectivity_hyp=U(0.0, 0.2)),
        lambda: ns.MirrorCube(pos(), L=U(1.0, 2.5)),
        lambda: ns.MLA(pos(), N=(int(rng.integers(2, 5)), int(rng.integers(2, 4))), pitch=U(0.4, 0.8), focal_length=U(2, 6),
                       radius=U(0.15, 0.35)),
        lambda: ns.DMD(pos(), N=(int(rng.integers(2, 4)), int(rng.integers(2, 4))), pitch=U(0.4, 0.8), tilt_angle=U(0.5, 1.0)),
        lambda: ns.ASphericParametricLens(pos(), CT=U(0.5, 0.9), diameter=U(2.0, 3.5), n=1.0 + U(0.4, 0.7), R=U(4, 9),
                                          kappa=-U(0.2, 1.2), a4=U(-2e-4, 2e-4), a6=U(-5e-6, 5e-6)),
        lambda: ns.SphereRefractive(pos(), radius=U(2.0, 6.0), height=U(0.5, 1.5), n1=1.0, n2=1.0 + U(0.3, 0.8),
                                    reflectivity=U(0.0, 0.15)),
    ]
    if extended:
        # the rest of the zoo (CPU-side tests only: the GPU fuzz tests keep the scenes their seeds have always meant)
        palette += [
            lambda: ns.MMA(pos(), N=(int(rng.integers(2, 5)), int(rng.integers(1, 4))), pitch=U(0.4, 0.8), roc=U(3, 12), n=1.0 + U(0.4, 0.6),
                           thickness=U(0.2, 0.6)),
            lambda: ns.MirrorPair(pos(), width=U(1.5, 3), height=U(1.5, 3), angle=U(1.2, 1.9), reflectivity_1=U(0.5, 1.0),
                                  transmission_2=U(0.0, 0.4)),
            lambda: ns.MirrorPrism(pos(), width=U(1.5, 3), height=U(1.5, 3), angle=U(1.2, 1.9), reflectivity=U(0.6, 1.0)),
            lambda: ns.TriangularPrism(pos(), width=U(1.5, 3), height=U(1.5, 3), n1=1.0, n2=ns.Glass_NBK7(), alpha=U(0.6, 1.0),
                                       beta=U(1.2, 1.7), reflectivity_1=U(0.0, 0.3), reflectivity_2=U(0.0, 0.5),
                                       max_interact_count_2=int(rng.integers(2, 6)), max_interact_count_3=int(rng.integers(2, 6))),
            lambda: ns.DovePrism(pos(), L=U(3, 5), D=U(0.8, 1.2), Ng=1.0 + U(0.4, 0.6)),
            lambda: ns.ASphericExactSphericalLens(pos(), EFL=U(6, 15), CT=U(0.5, 0.9), diameter=U(2.0, 3.0), n=1.0 + U(0.4, 0.7)),
            lambda: ns.SquareRefractive(pos(), width=U(1.5, 4), height=U(1.5, 4), n1=1.0, n2=1.0 + U(0.3, 0.8), reflectivity=U(0.0, 0.3)),
            lambda: ns.CircleRefractive(pos(), radius=U(0.8, 2.0), n1=1.0 + U(0.0, 0.5), n2=1.0 + U(0.3, 0.8), reflectivity=U(0.0, 0.3)),
        ]
    comps = []
    for k in rng.choice(len(palette), size=int(rng.integers(4, 10))):
        c = place(palette[int(k)]())
        if U() < 0.15:
            c.max_interact_count = None  # (exercise the attribute without binding caps: family-serial has its own tests)
        comps.append(c)
    wls = (450e-7, 633e-7, 1064e-7)
    rays = []
    for k in range(n_rays):
        kw = {"wavelength": wls[k % 3]}
        if k % 4:
            kw["w0"] = U(20e-4, 80e-4)
        if k % 11 == 5:
            kw["length"] = U(4, 40)
        d = [1.0, 0.12 * rng.standard_normal(), 0.04 * rng.standard_normal()]
        rays.append(ns.Ray([U(-1, 1), U(-1.5, 1.5), U(-0.5, 0.5)], d, intensity=U(0.2, 1.0), **kw))
    mons = [place(ns.Monitor(pos(), U(4, 12), U(4, 12))), ns.Monitor([20, 0, 0], 30, 12)]
    if caps:
        # binding interact caps: a third of the leaves stop interacting after 1-3 hits per ray id, and
        # the rays come in families of three wavelengths sharing one id, so the outcome depends on the reference's
        # sequential order across the rays of a family and across pops
        def leaves(cs, out):
            for c in cs:
                leaves(c.components, out) if hasattr(c, "components") else out.append(c)
            return out

        for leaf in leaves(comps, []):
            if U() < 0.33:
                leaf.max_interact_count = int(rng.integers(1, 4))
        fam = []
        for k, r in enumerate(rays[: max(3, n_rays // 3)]):
            fam += [r.copy(wavelength=w, id=1000 + k) if hasattr(r, "copy") else r for w in wls]
        rays = fam
    return Scene(comps, rays, mons, limit={"max_trace_num": 60})


def callable_material(ns):
    """a19: `Material(name, n=<arbitrary Python callable>)` (material.py:4-21) on both sides of a tilted slab and in a
    plano-convex lens, three wavelengths sharing ray ids (multiplex_rays_in_wavelength): a Cauchy law written as a
    plain lambda, which the device can only get as a per-wavelength table evaluated by the host."""
    cauchy = ns.Material("Cauchy", n=lambda wl_m: 1.5046 + 4.2e-15 / wl_m ** 2)
    flint = ns.Material("tab", n=lambda wl_m: float(np.interp(wl_m, [3e-7, 6e-7, 12e-7], [1.74, 1.68, 1.65])))
    slab = ns.GlassSlab([4, 0, 0], width=4, height=4, thickness=0.7, n1=ns.Vacuum(), n2=cauchy, reflectivity=0.05).RotZ(0.3)
    face = ns.CircleRefractive([8, 0, 0], radius=2.0, n1=1.0, n2=flint)
    back = ns.SphereRefractive([8.9 - 6.0, 0, 0], radius=6.0, height=0.4, n1=1.0, n2=flint)
    mon = ns.Monitor([14, 0, 0], 8, 8)
    rng = np.random.default_rng(SEED + 9)
    r0 = [ns.Ray([0, y, z], [1, 0.02 * rng.standard_normal(), 0.02 * rng.standard_normal()], wavelength=780e-7, w0=50e-4)
          for y, z in rng.uniform(-0.8, 0.8, (5, 2))]
    rays = ns.multiplex_rays_in_wavelength(r0, [780e-7, 532e-7, 405e-7])
    return Scene([slab, face, back], rays, [mon], limit={"max_trace_num": 60})


def stale_boxes(ns):
    """`.bbox` caches on first use (optical_component.py:63-67, component_group.py:28-32) and a later move
    does not refresh it, so ComponentGroup.interact (component_group.py:98-107) tests boxes of where things WERE.
    `near` sits at x = 5 but its box still says x = 20: the reference finds it (the ray crosses the old box too) IN FRONT
    of the splitter at x = 10, so its box must not be used to dismiss it by distance. `ghost` was moved into the beam at
    x = 15 but its box is still at y = 30: the reference never sees it. `lens` moved along the beam with a fresh box."""
    near = ns.CircleRefractive([20, 0, 0], radius=2.0, n1=1.0, n2=1.5)
    ghost = ns.Mirror([15, 30, 0], radius=2.0)
    _ = near.bbox, ghost.bbox
    near.TX(-15)
    ghost.TY(-30)
    split = ns.BeamSplitter([10, 0, 0], width=4, height=4, eta=0.4).RotZ(0.2)
    back = ns.SphereRefractive([27.0 - 9.0, 0, 0], radius=9.0, height=0.3, n1=1.5, n2=1.0)
    g = ns.ComponentGroup([0, 0, 0])
    g.add_components([split, near, ghost, back])
    _ = g.bbox
    lens = ns.Doublet([40, 0, 0], CT1=1.359, CT2=0.6, R1=18.405, R2=-13.734, R3=-39.933, n12=ns.Glass_NBK7(),
                      n23=ns.Glass_NSF5(), diameter=7.5)
    _ = lens.bbox
    lens.TX(-6)                      # children moved with it; every cached box (group and children) is stale now
    mon = ns.Monitor([60, 0, 0], 12, 12)
    rng = np.random.default_rng(SEED + 10)
    rays = [ns.Ray([0, y, z], [1, 0.01 * rng.standard_normal(), 0.01 * rng.standard_normal()], wavelength=633e-7, w0=50e-4)
            for y, z in rng.uniform(-1.2, 1.2, (8, 2))]
    return Scene([g, lens], rays, [mon], limit={"max_trace_num": 60})


def monitor_as_component(ns):
    """`Monitor.interact_local` returns `[ray]` (monitor.py:174-175), so a Monitor listed in
    table.components is a pass-through that hands the popped ray back UNCHANGED -- same origin -- to be hit again at the
    same distance on every later pop, until the pop cap drops it. Degenerate, but it is what the reference does: rays that
    reach the screen stall there (25 pops each here), rays that miss it go on to the mirror; the same Monitor object
    listed in table.monitors records every one of those dead segments."""
    lens = ns.Lens([4, 0, 0], focal_length=9.0, radius=2.0)
    screen = ns.Monitor([8, 0.5, 0], 1.2, 1.2).RotZ(0.15)
    mirror = ns.Mirror([12, 0, 0], radius=3.0).RotZ(2.8)
    far = ns.Monitor([6, -6, 0], 6, 6).RotZ(-1.2)
    rng = np.random.default_rng(SEED + 11)
    rays = [ns.Ray([0, y, z], [1, 0.01 * rng.standard_normal(), 0.01 * rng.standard_normal()], wavelength=633e-7, w0=40e-4)
            for y, z in rng.uniform(-1.5, 1.5, (10, 2))]
    return Scene([lens, screen, mirror], rays, [screen, far], limit={"max_trace_num": 25})


REGISTRY = {
    "gaussian_beam": gaussian_beam,
    "glass_slab": glass_slab,
    "chromatic": chromatic,
    "cavity_misaligned": cavity,
    "cavity_aligned": cavity_aligned,
    "doublet": doublet,
    "telescope_4f": telescope_4f,
    "exact_asphere": exact_asphere,
    "mirror_pair": mirror_pair,
    "prism_refl": prism_refl,
    "dove_prism": dove_prism,
    "gaussian_telescope": gaussian_telescope,
    "misc_components": misc_components,
    "mma_small": mma_small,
    "caps_binding": caps_binding,
    "extras": extras,
    "ripa": lambda ns: ripa(ns, n_rays=3, limit=150),
    "ripa2_simplified": ripa2_simplified,
    "callable_material": callable_material,
    "stale_boxes": stale_boxes,
    "monitor_as_component": monitor_as_component,
    "nested_csg": nested_csg,
}
# a dozen random scenes are ordinary fixtures too (reference-generated goldens); the fuzz tests add hundreds more
for _seed in range(12):
    REGISTRY[f"fuzz_{_seed:02d}"] = (lambda ns, _s=_seed: fuzz(ns, _s))


# ---- array-level ray batches for the scale tests (no Python Ray objects) ---------------------------------------
def ray_arrays(n, origin, spread_pos, direction, spread_dir, wavelengths=(780e-7,), w0=61e-4, seed=SEED + 100):
    """n rays: origin + U(-spread_pos, spread_pos) per axis, direction + N(0, spread_dir) (normalised), wavelengths
    cycled; Gaussian q from w0 (0 = no q). Returns the dict layout of optable_b200.flatten.pack_rays."""
    rng = np.random.default_rng(seed)
    o = np.asarray(origin, float) + rng.uniform(-1, 1, (n, 3)) * np.asarray(spread_pos, float)
    d = np.asarray(direction, float) + rng.standard_normal((n, 3)) * np.asarray(spread_dir, float)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    wl = np.asarray(wavelengths, float)[np.arange(n) % len(wavelengths)]
    arrs = {"ox": o[:, 0].copy(), "oy": o[:, 1].copy(), "oz": o[:, 2].copy(),
            "dx": d[:, 0].copy(), "dy": d[:, 1].copy(), "dz": d[:, 2].copy(),
            "intensity": np.ones(n), "wavelength": wl, "pathlength": np.zeros(n), "n_medium": np.ones(n),
            "length": np.full(n, np.inf)}
    if w0:
        arrs["q_re"], arrs["q_im"] = np.zeros(n), np.pi * w0 ** 2 / wl
    else:
        arrs["q_re"], arrs["q_im"] = np.zeros(n), np.zeros(n)
    arrs["flags"] = np.full(n, 1 | (2 if w0 else 0), dtype=np.uint32)
    arrs["family"] = np.arange(n, dtype=np.int32)
    return arrs
