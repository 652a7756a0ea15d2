"""Monitor analytics on the device: the reference's `Monitor` accessors (optable/monitor.py:53-269) evaluated on
the row columns a bundle trace leaves in HBM, without copying rows to the host or building Python tuples.

`trace_bundle` returns one set of row columns for all monitors; `DeviceMonitor(table.monitors[m], out, m)` selects
monitor m's rows and offers the same quantities as tensors on the same device. PyTorch is the array library here
(sorting, reductions); the rows themselves come from the CUDA path. Row order "YZ" = lexicographic (y, z) of the
monitor-local point like `Monitor.sortYZIndex`; "ID" = by initial ray (a bundle ray's id is its index); None =
device append order.
"""
from __future__ import annotations

import numpy as np


class DeviceMonitor:
    def __init__(self, monitor, out: dict, index: int):
        import torch

        self.monitor, self.torch = monitor, torch
        sel = torch.nonzero(out["hit_monitor"] == index).squeeze(1)
        col = lambda k: out[k][sel] if k in out else None
        self.root, self.pop = col("hit_root"), col("hit_pop")
        self.P = torch.stack([col("hit_px"), col("hit_py"), col("hit_pz")], 1)
        self.I, self.t = col("hit_intensity"), col("hit_t")
        d = [col("hit_dx"), col("hit_dy"), col("hit_dz")]
        self.direction = torch.stack(d, 1) if d[0] is not None else None
        qr, qi = col("hit_q_re"), col("hit_q_im")
        self.q = torch.complex(qr, qi) if qr is not None else None
        dev, f64 = self.P.device, torch.float64
        self._tY = torch.as_tensor(np.asarray(monitor.tangent_Y, dtype=np.float64), device=dev, dtype=f64)
        self._tZ = torch.as_tensor(np.asarray(monitor.tangent_Z, dtype=np.float64), device=dev, dtype=f64)
        self._normal = torch.as_tensor(np.asarray(monitor.normal, dtype=np.float64), device=dev, dtype=f64)
        self.hist_y = out["hist_y"][index] if out.get("hist_y") is not None else None
        self._orders = {}

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
        return self.get_PList(sort) @ self._tY   # monitor-LOCAL point . LAB tangent: the reference's quirk, kept

    def get_zList(self, sort="YZ"):
        return self.get_PList(sort) @ self._tZ

    def get_IList(self, sort="YZ"):
        return self.I[self.order(sort)]

    def get_tList(self, sort="YZ"):
        return self.t[self.order(sort)]

    def get_directionList(self, sort="YZ"):
        return self.direction[self.order(sort)]

    def get_tYList(self, sort="YZ"):
        return self.get_directionList(sort) @ self._tY

    def get_tZList(self, sort="YZ"):
        return self.get_directionList(sort) @ self._tZ

    @property
    def sum_intensity(self):
        return self.I.sum()

    @property
    def avg_intensity(self):
        return self.I.mean()

    def get_waist_distance(self, sort="YZ"):
        """Distance from every hit to its beam waist, signed along the monitor normal (monitor.py:202-216)."""
        o = self.order(sort)
        z = (self.q[o] + self.t[o]).real
        return self.torch.where(self.direction[o] @ self._normal > 0, -z, z)

    def get_delta_pos(self):
        y, z = self.get_yList(), self.get_zList()
        if y.numel() == 0:
            zero = self.torch.zeros(1, dtype=self.torch.float64, device=self.P.device)
            return zero, zero.clone()
        idx = self.torch.sort(y, stable=True).indices
        return self.torch.diff(y[idx]), self.torch.diff(z[idx])

    def _get_hist_y(self):
        """(counts, bin edges) of the 30-bin histogram of yList over +-width/2 (monitor.py:195-200): the histogram
        the trace accumulated when it ran with record_hist, else computed from the rows with the same binning."""
        torch = self.torch
        w = float(self.monitor.width)
        edges = torch.linspace(-w / 2, w / 2, 31, dtype=torch.float64, device=self.P.device)
        if self.hist_y is not None and int(self.hist_y.sum()) > 0:
            return self.hist_y, edges
        y = self.get_yList(sort=None)
        inside = (y >= -w / 2) & (y <= w / 2)
        b = torch.clamp(torch.bucketize(y[inside], edges, right=True) - 1, max=29)   # last bin closed, like numpy
        return torch.bincount(b, minlength=30), edges

    @property
    def std_histy(self):
        counts, edges = self._get_hist_y()
        c = counts.to(self.torch.float64)
        left = edges[:-1]
        mean = (c * left).sum() / c.sum()
        return self.torch.sqrt((c * left ** 2).sum() / c.sum() - mean ** 2)

    def export_rays_npz(self, filename: str):
        """Same arrays as Monitor.export_rays_npz (monitor.py:255-269); this is the one place rows leave the device."""
        print(f"Exporting {self.ndata} rays to {filename} ...")
        host = lambda x: x.cpu().numpy()
        np.savez(filename, xList=host(self.get_yList()), yList=host(self.get_zList()), tXList=host(self.get_tYList()),
                 tYList=host(self.get_tZList()), IList=host(self.get_IList()))
