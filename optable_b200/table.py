"""OpticalTable: the scene container and THE drop-in boundary (reference: optable/optical_table.py:9-147).

`OpticalTable.ray_tracing(rays, perfomance_limit=None)` keeps the reference's signature and observable effects
(SURVEY 8b): `self.rays` is extended with every traced segment in (initial ray, pop) order, the return value is
a copy of `self.rays`, every monitor receives its `(P_local, intensity, t, ray)` rows, capped components get their
`_interact_count` updated, inputs are not mutated. The work itself is one call into liboptb.so.

`trace_table` is duck-typed, so `install(reference_module)` can also swap the back end under the reference's own
classes. `trace_bundle` is the tensor entry point for batches too large to exist as Python objects.
"""
from __future__ import annotations

import copy as _copy
import math
from typing import List, Union

import numpy as np

from . import _abi as A
from .assemblies import ComponentGroup
from .elements import OpticalComponent
from .flatten import FlatScene, batch_wavelengths_m, pack_rays, trace_cap
from .monitors import Monitor
from .rays import Ray


_CONST_MATERIALS = {}


def _const_material(n):
    """One shared constant-index Material per distinct value (what `Ray._n = n` would build for every segment)."""
    m = _CONST_MATERIALS.get(n)
    if m is None:
        from .materials import Material

        if len(_CONST_MATERIALS) > 4096:
            _CONST_MATERIALS.clear()
        m = _CONST_MATERIALS[n] = Material("Constant", n=n)
    return m


def _segment_objects(rays, out):
    """Segment rows -> ray objects: a shallow copy of the initial ray (keeps `_id`, wavelength, user attributes)
    with the traced fields overwritten. Columns are converted to Python scalars in bulk; for this package's own
    `Ray` the copy is a `__dict__` update instead of the generic copy protocol."""
    n = len(out["seg_root"])
    if not n:
        return []
    O = np.stack((out["seg_ox"], out["seg_oy"], out["seg_oz"]), 1)
    D = np.stack((out["seg_dx"], out["seg_dy"], out["seg_dz"]), 1)
    flags = out["seg_flags"]
    root = out["seg_root"].tolist()
    inf = np.isinf(out["seg_length"])
    length = out["seg_length"].tolist()
    alive = ((flags & A.RF_ALIVE) != 0).tolist()
    hasq = ((flags & A.RF_HASQ) != 0).tolist()
    inten = out["seg_intensity"].tolist()
    q = (out["seg_q_re"] + 1j * out["seg_q_im"]).tolist()
    pl = out["seg_pathlength"].tolist()
    nmed = out["seg_n"].tolist()
    if inf.any():
        for k in np.nonzero(inf)[0].tolist():
            length[k] = None
    segs = [None] * n
    new = object.__new__
    for k in range(n):
        src = rays[root[k]]
        if type(src) is Ray:
            seg = new(Ray)
            d = seg.__dict__
            d.update(src.__dict__)
            d["origin"] = O[k].copy()
            d["_direction"] = D[k].copy()
            d["length"] = length[k]
            d["alive"] = alive[k]
            d["intensity"] = inten[k]
            d["qo"] = q[k] if hasq[k] else None
            d["_pathlength"] = pl[k]
            d["_n"] = _const_material(nmed[k])
        else:  # any other ray class (the reference's own after `install`): generic copy + attribute protocol
            seg = _copy.copy(src)
            seg.origin = O[k].copy()
            seg._direction = D[k].copy()
            seg.length = length[k]
            seg.alive = alive[k]
            seg.intensity = inten[k]
            seg.qo = q[k] if hasq[k] else None
            seg._pathlength = pl[k]
            seg._n = nmed[k]
        segs[k] = seg
    return segs


def trace_table(table, rays, perfomance_limit=None, engine=None):
    """Trace `rays` through `table.components`, fill `table.monitors`, return the new segment objects.
    Works on any objects shaped like the reference's (attributes only)."""
    from .backend import Engine

    engine = engine or Engine.get()
    arrs, fam_ids, unit = pack_rays(rays)
    flat = FlatScene(table.components, table.monitors, wavelengths_m=batch_wavelengths_m(arrs["wavelength"], unit))
    caps = None
    if flat.n_capslots:
        caps = np.zeros((flat.n_capslots, len(fam_ids)), np.int32)
        for s, comp in enumerate(flat.capslots):
            for f, rid in enumerate(fam_ids):
                caps[s, f] = comp._interact_count.get(rid, 0)
    scene = engine.upload(flat)
    try:
        out = engine.trace_arrays(scene, arrs, max_trace_num=trace_cap(perfomance_limit), unit=unit,
                                  n_families=len(fam_ids), cap_counts=caps)
    finally:
        scene.close()
    if int(out["counters"][A.C_STATUS]) & A.ST_CAP_ORDER:
        raise RuntimeError("an interact cap (max_interact_count) bound while several rays of one family were in "
                           "flight: the reference result depends on sequential order; not supported on the device")
    segs = _segment_objects(rays, out)
    # segment lookup for monitor rows: rows are (root, pop)-sorted, so is the segment list
    if len(out["hit_root"]):
        seg_key = out["seg_root"].astype(np.int64) << 32 | out["seg_pop"].astype(np.int64)
        hit_key = out["hit_root"].astype(np.int64) << 32 | out["hit_pop"].astype(np.int64)
        seg_of_hit = np.searchsorted(seg_key, hit_key)
        for mi, mon in enumerate(table.monitors):
            rows = np.nonzero(out["hit_monitor"] == mi)[0]
            if not len(rows):
                continue
            P = np.stack([out["hit_px"][rows], out["hit_py"][rows], out["hit_pz"][rows]], 1)
            robjs = [segs[j] for j in seg_of_hit[rows]]
            if hasattr(mon, "_extend"):
                mon._extend(P, out["hit_intensity"][rows], out["hit_t"][rows],
                            np.stack([out["hit_dx"][rows], out["hit_dy"][rows], out["hit_dz"][rows]], 1),
                            out["hit_q_re"][rows] + 1j * out["hit_q_im"][rows], [r._id for r in robjs], robjs)
            else:  # a reference-class Monitor: same tuples Monitor.record appends (monitor.py:188-193)
                mon._data_raw.extend((P[i], float(out["hit_intensity"][rows[i]]), float(out["hit_t"][rows[i]]), robjs[i])
                                     for i in range(len(rows)))
                mon._updated = True
    if flat.n_capslots:
        for s, comp in enumerate(flat.capslots):
            for f, rid in enumerate(fam_ids):
                if out["cap_counts"][s, f]:
                    comp._interact_count[rid] = int(out["cap_counts"][s, f])
    return segs


def single_pop(component, ray, engine=None):
    """`component.interact(ray)` of the reference (optical_component.py:337-378, component_group.py:93-122) for one
    component or group: one pop of the bounce loop against this component alone, on the device.
    Returns (t, [truncated parent, child rays...]) or (None, None). Interact counts of capped leaves are updated
    for this pop only, like the reference's.

    Two device calls: with a pop budget of 1 the children are queued and dropped, which yields t, the count update
    and (from the dropped counter) the number of children; a second call with budget 1 + children, on a scratch
    count table, pops the children so that their state can be read from their segment rows."""
    from .backend import Engine

    if not ray.alive:
        return None, None
    engine = engine or Engine.get()
    arrs, fam_ids, unit = pack_rays([ray])
    flat = FlatScene([component], [], wavelengths_m=batch_wavelengths_m(arrs["wavelength"], unit))
    caps = None
    if flat.n_capslots:
        caps = np.array([[c._interact_count.get(ray._id, 0)] for c in flat.capslots], np.int32)
    scene = engine.upload(flat)
    try:
        first = engine.trace_arrays(scene, arrs, max_trace_num=1, unit=unit, n_families=1, cap_counts=caps)
        if int(first["seg_leaf"][0]) < 0:
            return None, None
        if flat.n_capslots:
            for s, comp in enumerate(flat.capslots):
                if first["cap_counts"][s, 0]:
                    comp._interact_count[ray._id] = int(first["cap_counts"][s, 0])
        t = float(first["seg_length"][0])
        n_children = int(first["counters"][A.C_DROPPED])
        out = [ray.copy(length=t, alive=False)]
        if n_children:
            again = engine.trace_arrays(scene, arrs, max_trace_num=1 + n_children, unit=unit, n_families=1, cap_counts=caps)
            for k in range(1, 1 + n_children):
                hasq = bool(again["seg_flags"][k] & A.RF_HASQ)
                child = ray.copy(alive=True)
                child.origin = np.array([again["seg_ox"][k], again["seg_oy"][k], again["seg_oz"][k]])
                child._direction = np.array([again["seg_dx"][k], again["seg_dy"][k], again["seg_dz"][k]])
                child.intensity = float(again["seg_intensity"][k])
                child.qo = complex(again["seg_q_re"][k], again["seg_q_im"][k]) if hasq else None
                child._pathlength = float(again["seg_pathlength"][k])
                child._n = float(again["seg_n"][k])
                out.append(child)
    finally:
        scene.close()
    return t, out


class OpticalTable:
    def __init__(self, **kwargs):
        self.components = []
        self.rays = []
        self.monitors = []
        self.norender_set = set()
        self._bbox = (None,) * 6
        self.unit = kwargs.get("unit", 1e-2)

    @staticmethod
    def _collect(target, item, kind):
        if isinstance(item, kind):
            target.append(item)
        elif isinstance(item, list):
            for entry in item:
                if isinstance(entry, kind):
                    target.append(entry)
                elif isinstance(entry, list):
                    target.extend(entry)  # one level of nesting, like the reference

    def add_components(self, component: Union[OpticalComponent, List]):
        self._collect(self.components, component, OpticalComponent)

    def add_monitors(self, monitor: Union[Monitor, List]):
        self._collect(self.monitors, monitor, Monitor)

    @property
    def bbox(self):
        if self._bbox[0] is None:
            self.get_bbox()
        return self._bbox

    def get_bbox(self):
        from .shapes import Surface

        self._bbox = Surface.merge_bboxs([c.bbox for c in self.components])
        return self._bbox

    def ray_tracing(self, rays: Union[Ray, List[Ray]], perfomance_limit=None):
        """Trace on the GPU; same contract as the reference method (keyword spelling included)."""
        if isinstance(rays, Ray):
            rays = [rays]
        self.rays.extend(trace_table(self, list(rays), perfomance_limit))
        return [r.copy() for r in self.rays]

    # ---- callers of ray_tracing that the reference ships on the table (optical_table.py:211-422) ----------
    def calculate_abcd_matrix(self, mon0: Monitor, mon1: Monitor, rays: List[Ray], disp=1e-5, rot=1e-5, debugaxs=None):
        """Per-ray 2x2 ray-transfer matrix from `mon0` to `mon1` along the monitors' Y axis, by finite differences:
        one nominal trace, one with every ray shifted by `disp` along mon0's tangent_Y, one with every (already
        shifted) ray additionally turned by `rot` about mon0's tangent_Z through its recorded mon0 hit point.
        Rays need distinct ids and must reach both monitors. Like the reference, `self.rays` and both monitors are
        cleared before each trace, the shift is not undone before the rotation, and the pivot is the hit point in
        monitor-local coordinates (optical_table.py:264-286). Returns an (N, 2, 2) array ordered by ray id."""
        if debugaxs is not None:
            raise NotImplementedError("rendering is out of scope of optable_b200")
        assert len(rays) > 0, "No rays to trace in ABCD calculation."
        ids = [r._id for r in rays]
        assert len(set(ids)) == len(rays), "Redundant ray ids in ABCD calculation."
        work = [rays[i].copy() for i in np.argsort(ids)]

        def run(batch):
            self.rays = []
            mon0.clear()
            mon1.clear()
            self.ray_tracing(batch)
            return mon1.get_yList(sort="ID"), mon1.get_tYList(sort="ID")

        shift, axis = mon0.tangent_Y * disp, mon0.tangent_Z
        both = self._abcd_first_two_traces(mon0, mon1, work, shift)
        if both is not None:
            # nominal and shifted bundle traced as ONE batch, rows read as columns (no segment objects); only the
            # last trace below goes through ray_tracing and leaves table.rays / the monitors as the reference does
            (y0, t0, pivots), (y1, t1, _) = both
            for r in work:
                r._Translate(shift)
        else:
            y0, t0 = run(work)
            for mon, label in ((mon0, "mon0"), (mon1, "mon1")):
                assert set(ids) == {r._id for r in mon.get_rays(sort="ID")}, f"Rays at {label} do not match the input rays."
            pivots = mon0.get_PList(sort="ID")
            y1, t1 = run([r._Translate(shift) for r in work])
        y2, t2 = run([r._RotAround(axis, pivots[k], rot) for k, r in enumerate(work)])
        Ms = np.zeros((len(rays), 2, 2))
        Ms[:, 0, 0], Ms[:, 1, 0] = (y1 - y0) / disp, (t1 - t0) / disp
        Ms[:, 0, 1], Ms[:, 1, 1] = (y2 - y0) / rot, (t2 - t0) / rot
        return Ms

    def _abcd_first_two_traces(self, mon0, mon1, work, shift):
        """Nominal + shifted bundle of calculate_abcd_matrix in one device call. Returns
        ((y, tY, pivots) nominal, (y, tY, pivots) shifted), rows ordered like `work` (= by ray id), or None when
        the shortcut would not be equivalent to three separate ray_tracing calls: other monitors on the table
        (they would miss the rows of these traces), interact caps (counts), or a subclass overriding ray_tracing."""
        from .backend import Engine

        mine = [m for m in self.monitors if m is mon0 or m is mon1]
        if type(self).ray_tracing is not OpticalTable.ray_tracing or len(mine) != len(self.monitors) or len(mine) != 2:
            return None
        n = len(work)
        batch = work + [r.copy()._Translate(shift) for r in work]
        arrs, fam_ids, unit = pack_rays(batch)
        flat = FlatScene(self.components, [mon0, mon1], wavelengths_m=batch_wavelengths_m(arrs["wavelength"], unit))
        if flat.n_capslots:
            return None
        engine = Engine.get()
        scene = engine.upload(flat)
        try:
            out = engine.trace_arrays(scene, arrs, max_trace_num=trace_cap(None), unit=unit, record_segments=False,
                                      n_families=len(fam_ids))
        finally:
            scene.close()
        P = np.stack([out["hit_px"], out["hit_py"], out["hit_pz"]], 1)
        D = np.stack([out["hit_dx"], out["hit_dy"], out["hit_dz"]], 1)
        root, mon = out["hit_root"].astype(np.int64), out["hit_monitor"]
        halves = []
        for half in (0, 1):
            rows0 = np.nonzero((mon == 0) & (root // n == half))[0]
            rows1 = np.nonzero((mon == 1) & (root // n == half))[0]
            for rows, label in ((rows0, "mon0"), (rows1, "mon1")):
                assert set((root[rows] % n).tolist()) == set(range(n)), f"Rays at {label} do not match the input rays."
            rows0 = rows0[np.argsort(root[rows0], kind="stable")]
            rows1 = rows1[np.argsort(root[rows1], kind="stable")]
            # (monitor-local point dotted with the lab tangent: the reference's get_yList, monitor.py:78-84)
            halves.append((P[rows1] @ mon1.tangent_Y, D[rows1] @ mon1.tangent_Y, P[rows0]))
        return halves

    @staticmethod
    def calibrate_symmetric_4f(lens, rays: List[Ray], F10: float, F20: float, criterion: str = "M=-I", debugaxs=None,
                               optimize=True, display_M=False):
        """Symmetric 4f relay mon0 - F1 - lens - 2 F2 - lens (turned by pi) - F1 - mon1 built from two copies of
        `lens`. With optimize=False: returns (Ms, yList, tYList) at (F10, F20). With optimize=True: Nelder-Mead
        over (F1, F2) (xatol 1e-5, 50 iterations) on the chosen criterion -- "M=-I": mean |M + I|, "flat_field":
        mean |d_s - d0| for d0 = 1.5, "min_stdtY": spread of the exit slopes -- and returns (F1, F2)
        (optical_table.py:299-422)."""
        if debugaxs is not None:
            raise NotImplementedError("rendering is out of scope of optable_b200")

        def simulate(F1, F2):
            first = lens.copy()._Translate(np.array([F1, 0, 0]) - lens.origin)
            second = lens.copy()._Translate(np.array([F1 + 2 * F2, 0, 0]) - lens.origin).RotZ(np.pi)
            m0, m1 = Monitor(origin=[0, 0, 0], width=5, height=5), Monitor(origin=[2 * F1 + 2 * F2, 0, 0], width=5, height=5)
            table = OpticalTable()
            table.add_components([first, second])
            table.add_monitors([m0, m1])
            table.ray_tracing(rays)
            y, ty = m1.get_yList(), m1.get_tYList()
            Ms = table.calculate_abcd_matrix(m0, m1, rays)
            if display_M:
                for M in Ms:
                    print(M)
            return Ms, y, ty

        def cost(F1, F2):
            Ms, _, ty = simulate(F1, F2)
            if criterion == "M=-I":
                return float(np.mean([np.linalg.norm(M + np.eye(2)) for M in Ms]))
            if criterion == "flat_field":
                d0 = 1.5
                return float(np.mean([abs((d0 * M[0, 0] - M[0, 1]) / (M[1, 1] - M[1, 0] * d0) - d0) for M in Ms]))
            if criterion == "min_stdtY":
                return float(np.std(ty))
            raise ValueError(f"Unknown criterion: {criterion}")

        if not optimize:
            return simulate(F10, F20)
        from scipy.optimize import minimize

        res = minimize(lambda x: cost(x[0], x[1]), x0=[F10, F20], method="Nelder-Mead",
                       options={"xatol": 1e-5, "maxiter": 50})
        return tuple(res.x)

    # ---- exports (optical_table.py:447-523) ----------------------------------------------------------
    def gather_rays_csv(self):
        from .export import ray_rows

        return ray_rows(self.rays)

    def gather_components(self, avoid_flatten_classname: List = [], ignore_classname: List = []) -> List[dict]:
        from .export import component_rows

        rows = []
        for c in self.components:
            rows.extend(component_rows(c, avoid_flatten_classname, ignore_classname))
        return rows

    def export_rays_csv(self, filename: str):
        from .export import write_csv

        print(f"Exporting rays to {filename} ...")
        write_csv(filename, self.gather_rays_csv())

    def export_components_csv(self, filename: str, avoid_flatten_classname: List = [], ignore_classname: List = []):
        from .export import write_csv

        print(f"Exporting components to {filename} ...")
        write_csv(filename, self.gather_components(avoid_flatten_classname, ignore_classname))

    def trace_bundle(self, bundle, perfomance_limit=None, group=None, gather=False, **kw):
        """Tensor entry point: see optable_b200.bundle.trace_bundle. Under an initialised torch.distributed process
        group (one rank per GPU) pass `group=` (or `group=True` for the default group): the bundle is sharded over
        the ranks and the monitors are merged (optable_b200.dist.trace_sharded)."""
        if group is not None and group is not False:
            from .dist import trace_sharded

            return trace_sharded(self, bundle, perfomance_limit, group=None if group is True else group, gather=gather, **kw)
        from .bundle import trace_bundle

        return trace_bundle(self, bundle, perfomance_limit, **kw)


def install(reference_module, engine=None):
    """Swap the back end under the reference package's own classes: `reference_module.OpticalTable.ray_tracing`
    becomes a call into liboptb.so (SURVEY 8b "install mechanism"). Returns the original method (assign it back to
    undo). `engine` defaults to the CUDA engine of the current device."""
    original = reference_module.OpticalTable.ray_tracing
    ray_cls = reference_module.Ray

    def ray_tracing(self, rays, perfomance_limit=None):
        if isinstance(rays, ray_cls):
            rays = [rays]
        self.rays.extend(trace_table(self, list(rays), perfomance_limit, engine=engine))
        return _copy.deepcopy(self.rays)

    reference_module.OpticalTable.ray_tracing = ray_tracing
    return original
