"""Monitor: a rectangular, non-blocking detector plane (reference: optable/monitor.py:5-269).

The reference records hits in a post-pass over the dead segments (Monitor.record, monitor.py:183-193). Here
the device tests every segment against every monitor as it is produced and appends rows
(P_local, intensity, t, key of the segment); `OpticalTable.ray_tracing` hands them over through `_extend`.
Rows are kept as columns; the reference's list-of-tuples view `_data_raw` is materialised on demand, so the
accessors work unchanged while 1e7-row monitors stay cheap.

Known reference quirks kept on purpose: `get_yList`/`get_zList` dot the monitor-LOCAL point with
the LAB-frame tangents.
"""
from __future__ import annotations

import numpy as np

from .elements import OpticalComponent
from .shapes import Rectangle


class Monitor(OpticalComponent):
    def __init__(self, origin, width, height, **kwargs):
        super().__init__(origin, **kwargs)
        self.width, self.height = width, height
        self.surface = Rectangle(width, height)
        self._initialize()

    def _initialize(self):
        self._P = np.zeros((0, 3))
        self._I = np.zeros(0)
        self._t = np.zeros(0)
        self._dir = np.zeros((0, 3))
        self._q = np.zeros(0, dtype=complex)
        self._ids = []          # Ray._id per row (sort="ID")
        self._rays = []         # segment objects per row (None when the trace did not materialise segments)
        self._sorted_cache = {}
        self._updated = False
        self.hist_y = None      # device-accumulated histograms of the last trace (record_hist)
        self.hist_yz = None

    def clear(self):
        self._initialize()

    # -- filled by the back end -----------------------------------------------------------------------
    def _extend(self, P, I, t, direction, q, ids, rays=None):
        n = len(I)
        self._P = np.concatenate([self._P, np.asarray(P, float).reshape(n, 3)])
        self._I = np.concatenate([self._I, np.asarray(I, float)])
        self._t = np.concatenate([self._t, np.asarray(t, float)])
        self._dir = np.concatenate([self._dir, np.asarray(direction, float).reshape(n, 3)])
        self._q = np.concatenate([self._q, np.asarray(q, complex)])
        self._ids.extend(ids)
        self._rays.extend(rays if rays is not None else [None] * n)
        self._sorted_cache = {}
        self._updated = True

    def record(self, rays, engine=None):
        """Test an explicit list of rays/segments against this monitor and append the hits (monitor.py:183-193).
        `OpticalTable.ray_tracing` fills monitors during the trace; this entry point is for segments that come
        from somewhere else. It is the same device code: the rays are traced through a scene that holds no
        component, where every ray is popped once, hits nothing and is tested against the monitor over its own
        `length`. Rows reference the ray objects passed in, like the reference's."""
        from .backend import Engine
        from .flatten import FlatScene, pack_rays

        rays = list(rays)
        self._updated = True
        if not rays:
            return
        engine = engine or Engine.get()
        arrs, fam_ids, unit = pack_rays(rays)
        scene = engine.upload(FlatScene([], [self]))
        try:
            out = engine.trace_arrays(scene, arrs, max_trace_num=1, unit=unit, n_families=len(fam_ids))
        finally:
            scene.close()
        src = [rays[k] for k in out["hit_root"].tolist()]
        self._extend(np.stack([out["hit_px"], out["hit_py"], out["hit_pz"]], 1), out["hit_intensity"], out["hit_t"],
                     np.stack([out["hit_dx"], out["hit_dy"], out["hit_dz"]], 1),
                     out["hit_q_re"] + 1j * out["hit_q_im"], [r._id for r in src], src)

    def interact_local(self, ray):
        return [ray]  # a monitor never alters a ray (monitor.py:174-175)

    # -- the reference's views --------------------------------------------------------------------------
    @property
    def ndata(self):
        return len(self._I)

    @property
    def _data_raw(self):
        return [(self._P[i], float(self._I[i]), float(self._t[i]), self._rays[i]) for i in range(self.ndata)]

    def _order(self, sort):
        if sort not in self._sorted_cache:
            if sort == "YZ":
                idx = np.lexsort((self._P[:, 2], self._P[:, 1]))
            elif sort == "ID":
                idx = np.argsort(self._ids)
            else:
                idx = np.arange(self.ndata)
            self._sorted_cache[sort] = idx
        return self._sorted_cache[sort]

    @property
    def sortYZIndex(self):
        return self._order("YZ")

    @property
    def sortIDindex(self):
        return self._order("ID")

    def get_data(self, sort="YZ"):
        raw = self._data_raw
        self._updated = False
        return [raw[i] for i in self._order(sort)]

    data = property(lambda self: self.get_data())

    def get_rays(self, sort="YZ"):
        return [self._rays[i] for i in self._order(sort)] if self.ndata else []

    rays = property(lambda self: self.get_rays())

    @property
    def raw_yList(self):
        return self._P[:, 1].copy() if self.ndata else np.array([])

    @property
    def raw_zList(self):
        return self._P[:, 2].copy() if self.ndata else np.array([])

    def _col(self, values, sort):
        return values[self._order(sort)] if self.ndata else np.array([])

    def get_PList(self, sort="YZ"):
        return self._col(self._P, sort)

    def get_yList(self, sort="YZ"):
        return self._col(self._P, sort) @ self.tangent_Y if self.ndata else np.array([])

    def get_zList(self, sort="YZ"):
        return self._col(self._P, sort) @ self.tangent_Z if self.ndata else np.array([])

    def get_IList(self, sort="YZ"):
        return self._col(self._I, sort)

    def get_tList(self, sort="YZ"):
        return self._col(self._t, sort)

    def get_directionList(self, sort="YZ"):
        return self._col(self._dir, sort)

    def get_tYList(self, sort="YZ"):
        return self.get_directionList(sort=sort) @ self.tangent_Y

    def get_tZList(self, sort="YZ"):
        return self.get_directionList(sort=sort) @ self.tangent_Z

    PList = property(lambda self: self.get_PList())
    yList = property(lambda self: self.get_yList())
    zList = property(lambda self: self.get_zList())
    IList = property(lambda self: self.get_IList())
    tList = property(lambda self: self.get_tList())
    directionList = property(lambda self: self.get_directionList())
    tYList = property(lambda self: self.get_tYList())
    tZList = property(lambda self: self.get_tZList())

    def get_ray_i(self, idx):
        return self.rays[idx], [self.yList[idx], self.zList[idx], self.tYList[idx], self.tZList[idx], self.IList[idx]]

    def get_ray_id(self, ray_id):
        order = self._order("YZ")
        for k, i in enumerate(order):
            if self._ids[i] == ray_id:
                return self.get_ray_i(k)
        return None, None

    # -- analytics -----------------------------------------------------------------------------------
    def _get_hist_y(self):
        return np.histogram(self.yList, bins=30, range=(-self.width / 2, self.width / 2))

    def get_waist_distance(self):
        """Distance from each hit to its beam waist, signed along the monitor normal (monitor.py:202-216)."""
        q = self._col(self._q, "YZ") + self.get_tList()
        towards = self.get_directionList() @ self.normal > 0
        return np.where(towards, -np.real(q), np.real(q))

    def get_beam_waist(self):
        """Gaussian beam radius w at each hit, from q propagated to the hit (monitor.py:218-225; the reference
        reads a non-existent `rList` there and raises AttributeError, the intended quantity is returned here).
        Needs the segment objects (wavelength and index are per ray)."""
        out = []
        for r, t in zip(self.get_rays(), self.get_tList()):
            out.append(r.waist(r.q_at_z(t)))
        return np.array(out)

    def get_delta_pos(self):
        y, z = self.yList, self.zList
        if len(y) == 0:
            return np.array([0.0]), np.array([0.0])
        idx = np.argsort(y)
        return np.diff(y[idx]), np.diff(z[idx])

    @property
    def sum_intensity(self):
        return np.sum(self._I)

    @property
    def avg_intensity(self):
        return np.mean(self._I)

    @property
    def std_histy(self):
        counts, bins = self._get_hist_y()
        mean = np.sum(counts * bins[:-1]) / np.sum(counts)
        return np.sqrt(np.sum(counts * bins[:-1] ** 2) / np.sum(counts) - mean ** 2)

    def export_rays_npz(self, filename: str):
        print(f"Exporting {self.ndata} rays to {filename} ...")
        np.savez(filename, xList=self.yList, yList=self.zList, tXList=self.tYList, tYList=self.tZList, IList=self.IList)
