"""Monitor analytics on the device: the reference's `Monitor` accessors (optable/monitor.py:53-269) evaluated on
the row columns a bundle trace leaves in HBM, without copying rows to the host or building Python tuples.

`trace_bundle` returns one set of row columns for all monitors; `DeviceMonitor(table.monitors[m], out, m)` selects
monitor m's rows. The per-row quantities (y, z, slopes, waist distances) and every summary that needs no ordering
(row count, intensity sum / mean, moments and extrema of the spot, the 30-bin histogram behind `std_histy`) come from
ONE fused pass of the library's own kernel over the rows (`optb_monitor_stats`, include/optb.h). Only the accessors the
reference defines through a sort (`sort="YZ"` / `"ID"` views, `get_delta_pos`) order those columns with torch.sort
(PyTorch as the array library). Row order "YZ" = lexicographic (y, z) of the monitor-local point like
`Monitor.sortYZIndex`; "ID" = by initial ray (a bundle ray's id is its index); None = device append order.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi as A


class DeviceMonitor:
    def __init__(self, monitor, out: dict, index: int, engine=None):
        import torch

        from .backend import Engine, lib

        self.monitor, self.torch = monitor, torch
        any_col = out["hit_px"]
        dev = any_col.device
        self.engine = engine or Engine.get(dev.index or 0)
        n = int(any_col.numel())
        # ---- the fused pass: per-row y / z / tY / tZ / waist distance + all order-free summaries, straight from HBM
        res = A.Result()
        for k in A.HIT_I32 + A.HIT_U32 + A.HIT_F64 + ("hit_key",):
            if out.get(k) is not None:
                setattr(res, k, out[k].data_ptr())
        fr = A.MonitorFrame()
        fr.tangent_y[:] = [float(v) for v in monitor.tangent_Y]
        fr.tangent_z[:] = [float(v) for v in monitor.tangent_Z]
        fr.normal[:] = [float(v) for v in monitor.normal]
        fr.half_width, fr.half_height = float(monitor.width) / 2, float(monitor.height) / 2
        f64 = torch.float64
        has_d, has_q = out.get("hit_dx") is not None, out.get("hit_q_re") is not None
        cols = torch.empty((5, max(n, 1)), dtype=f64, device=dev)
        self._stats_dev = torch.empty(A.MS_STRIDE, dtype=f64, device=dev)
        L = lib()
        L.optb_monitor_stats.argtypes = [C.c_void_p, C.POINTER(A.Result), C.c_int64, C.c_int64, C.c_int, C.POINTER(A.MonitorFrame),
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        ptr = lambda r, ok=True: C.c_void_p(cols[r].data_ptr()) if ok else None
        st = torch.cuda.current_stream(dev).cuda_stream
        self.engine._check(L.optb_monitor_stats(self.engine._ctx, C.byref(res), 0, n, int(index), C.byref(fr),
                                                C.c_void_p(self._stats_dev.data_ptr()), ptr(0), ptr(1), ptr(2, has_d),
                                                ptr(3, has_d), ptr(4, has_d and has_q), C.c_void_p(st)))
        sel = torch.nonzero(~torch.isnan(cols[0, :n])).squeeze(1)   # rows of this monitor (the kernel marks the others NaN)
        col = lambda k: out[k][sel] if out.get(k) is not None else None
        self._y, self._z = cols[0, :n][sel], cols[1, :n][sel]
        self._ty = cols[2, :n][sel] if has_d else None
        self._tz = cols[3, :n][sel] if has_d else None
        self._wd = cols[4, :n][sel] if has_d and has_q else None
        if out.get("hit_key") is not None and out.get("hit_root") is None:
            key = out["hit_key"][sel]
            self.root, self.pop = (key >> 32) & 0xFFFFFFFF, key & 0xFFFFFF
        else:
            self.root, self.pop = col("hit_root"), col("hit_pop")
        self.P = torch.stack([col("hit_px"), col("hit_py"), col("hit_pz")], 1)
        self.I, self.t = col("hit_intensity"), col("hit_t")
        d = [col("hit_dx"), col("hit_dy"), col("hit_dz")]
        self.direction = torch.stack(d, 1) if d[0] is not None else None
        qr, qi = col("hit_q_re"), col("hit_q_im")
        self.q = torch.complex(qr, qi) if qr is not None and qi is not None else None
        self.hist_y = out["hist_y"][index] if out.get("hist_y") is not None else None
        self._orders = {}
        self._stats = None

    def stats(self) -> dict:
        """Order-free summaries of this monitor's rows from the fused kernel (one small device-to-host read)."""
        if self._stats is None:
            s = self._stats_dev.cpu().numpy()
            n = max(s[A.MS_COUNT], 1.0)
            my, mz = s[A.MS_SUM_Y] / n, s[A.MS_SUM_Z] / n
            self._stats = {
                "count": int(s[A.MS_COUNT]), "sum_intensity": float(s[A.MS_SUM_I]), "avg_intensity": float(s[A.MS_SUM_I] / n),
                "mean_y": float(my), "mean_z": float(mz),
                "std_y": float(np.sqrt(max(s[A.MS_SUM_YY] / n - my * my, 0.0))), "std_z": float(np.sqrt(max(s[A.MS_SUM_ZZ] / n - mz * mz, 0.0))),
                "min_y": float(s[A.MS_MIN_Y]), "max_y": float(s[A.MS_MAX_Y]), "min_z": float(s[A.MS_MIN_Z]), "max_z": float(s[A.MS_MAX_Z]),
                "mean_waist_distance": float(s[A.MS_SUM_WD] / n), "mean_tY": float(s[A.MS_SUM_TY] / n),
                "hist_y": s[A.MS_HIST:A.MS_HIST + A.HIST_BINS].astype(np.int64),
            }
        return self._stats

    @property
    def ndata(self) -> int:
        return int(self.I.numel())

    def order(self, sort="YZ"):
        """Row permutation (int64 tensor) for a sort key of the reference's accessors."""
        torch = self.torch
        if sort not in self._orders:
            if sort == "YZ":       # np.lexsort((z, y)): y primary, z secondary = stable sort by z, then by y
                by_z = torch.sort(self.P[:, 2], stable=True).indices
                idx = by_z[torch.sort(self.P[by_z, 1], stable=True).indices]
            elif sort == "ID":
                key = (self.root.to(torch.int64) & 0xFFFFFFFF) << 32 | (self.pop.to(torch.int64) & 0xFFFFFFFF)
                idx = torch.sort(key, stable=True).indices
            else:
                idx = torch.arange(self.ndata, device=self.P.device)
            self._orders[sort] = idx
        return self._orders[sort]

    # -- the accessors (tensors on the device) --------------------------------------------------------
    def get_PList(self, sort="YZ"):
        return self.P[self.order(sort)]

    def get_yList(self, sort="YZ"):
        return self._y[self.order(sort)]   # monitor-LOCAL point . LAB tangent: the reference's quirk, kept (kernel)

    def get_zList(self, sort="YZ"):
        return self._z[self.order(sort)]

    def get_IList(self, sort="YZ"):
        return self.I[self.order(sort)]

    def get_tList(self, sort="YZ"):
        return self.t[self.order(sort)]

    def get_directionList(self, sort="YZ"):
        return self.direction[self.order(sort)]

    def get_tYList(self, sort="YZ"):
        return self._ty[self.order(sort)]

    def get_tZList(self, sort="YZ"):
        return self._tz[self.order(sort)]

    @property
    def sum_intensity(self):
        return self.stats()["sum_intensity"]

    @property
    def avg_intensity(self):
        return self.stats()["avg_intensity"]

    def get_waist_distance(self, sort="YZ"):
        """Distance from every hit to its beam waist, signed along the monitor normal (monitor.py:202-216)."""
        return self._wd[self.order(sort)]

    def get_delta_pos(self):
        y, z = self.get_yList(), self.get_zList()
        if y.numel() == 0:
            zero = self.torch.zeros(1, dtype=self.torch.float64, device=self.P.device)
            return zero, zero.clone()
        idx = self.torch.sort(y, stable=True).indices
        return self.torch.diff(y[idx]), self.torch.diff(z[idx])

    def _get_hist_y(self):
        """(counts, bin edges) of the 30-bin histogram of yList over +-width/2 (monitor.py:195-200), accumulated by
        the fused kernel with np.histogram's binning."""
        w = float(self.monitor.width)
        return self.stats()["hist_y"], np.linspace(-w / 2, w / 2, 31)

    @property
    def std_histy(self):
        counts, edges = self._get_hist_y()
        c = counts.astype(np.float64)
        left = edges[:-1]
        mean = (c * left).sum() / c.sum()
        return float(np.sqrt((c * left ** 2).sum() / c.sum() - mean ** 2))

    def export_rays_npz(self, filename: str):
        """Same arrays as Monitor.export_rays_npz (monitor.py:255-269); this is the one place rows leave the device."""
        print(f"Exporting {self.ndata} rays to {filename} ...")
        host = lambda x: x.cpu().numpy()
        np.savez(filename, xList=host(self.get_yList()), yList=host(self.get_zList()), tXList=host(self.get_tYList()),
                 tYList=host(self.get_tZList()), IList=host(self.get_IList()))
