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
from .flatten import FlatScene, pack_rays, trace_cap
from .monitors import Monitor
from .rays import Ray


def _segment_object(root_ray, o, d, length, alive, intensity, q, hasq, pathlength, n):
    seg = _copy.copy(root_ray)
    seg.origin = np.array(o, dtype=float)
    seg._direction = np.array(d, dtype=float)
    seg.length = None if math.isinf(length) else float(length)
    seg.alive = bool(alive)
    seg.intensity = float(intensity)
    seg.qo = complex(q) if hasq else None
    seg._pathlength = float(pathlength)
    seg._n = float(n)
    return seg


def trace_table(table, rays, perfomance_limit=None, engine=None):
    """Trace `rays` through `table.components`, fill `table.monitors`, return the new segment objects.
    Works on any objects shaped like the reference's (attributes only)."""
    from .backend import Engine

    engine = engine or Engine.get()
    flat = FlatScene(table.components, table.monitors)
    arrs, fam_ids, unit = pack_rays(rays)
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
    n = len(out["seg_root"])
    flags = out["seg_flags"]
    segs = [None] * n
    for k in range(n):
        segs[k] = _segment_object(
            rays[int(out["seg_root"][k])],
            (out["seg_ox"][k], out["seg_oy"][k], out["seg_oz"][k]), (out["seg_dx"][k], out["seg_dy"][k], out["seg_dz"][k]),
            out["seg_length"][k], flags[k] & A.RF_ALIVE, out["seg_intensity"][k],
            complex(out["seg_q_re"][k], out["seg_q_im"][k]), flags[k] & A.RF_HASQ, out["seg_pathlength"][k], out["seg_n"][k])
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

    def trace_bundle(self, bundle, perfomance_limit=None, **kw):
        """Tensor entry point: see optable_b200.bundle.trace_bundle."""
        from .bundle import trace_bundle

        return trace_bundle(self, bundle, perfomance_limit, **kw)


def install(reference_module):
    """Swap the back end under the reference package's own classes: `reference_module.OpticalTable.ray_tracing`
    becomes a call into liboptb.so (SURVEY 8b "install mechanism"). Returns the original method."""
    original = reference_module.OpticalTable.ray_tracing
    ray_cls = reference_module.Ray

    def ray_tracing(self, rays, perfomance_limit=None):
        if isinstance(rays, ray_cls):
            rays = [rays]
        self.rays.extend(trace_table(self, list(rays), perfomance_limit))
        return _copy.deepcopy(self.rays)

    reference_module.OpticalTable.ray_tracing = ray_tracing
    return original
