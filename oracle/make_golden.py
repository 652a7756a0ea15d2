"""Generate tests/golden/*.npz from the REAL reference (run in the build container only).

    python -m oracle.make_golden

Each fixture holds, for one scene of tests/scenes.py built with the reference's own classes:
the flattened tables (optable_b200.flatten.FlatScene.to_arrays), the packed input rays, the trace
parameters, and the reference's results (oracle.ref_harness.run_reference): every output segment in
(root, pop) order with the winning leaf index, and every Monitor._data_raw row.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_harness as RH  # noqa: E402
from optable_b200.flatten import FlatScene, pack_rays, trace_cap  # noqa: E402
from tests import scenes  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def make(name):
    ref = RH.load_reference()
    sc = scenes.REGISTRY[name](ref)
    flat = sc.flat()
    arrs, fam_ids, unit = pack_rays(sc.rays)
    r = RH.run_reference(sc)
    leaves = r.pop("_leaves")
    # interact counts after the trace, per cap slot and family
    caps = np.zeros((max(flat.n_capslots, 1), len(fam_ids)), np.int32)
    for s, comp in enumerate(flat.capslots):
        for f, rid in enumerate(fam_ids):
            caps[s, f] = comp._interact_count.get(rid, 0)
    d = {}
    d.update({"scene_" + k: v for k, v in flat.to_arrays().items()})
    d.update({"ray_" + k: v for k, v in arrs.items()})
    d["param_max_trace_num"] = np.int64(trace_cap(sc.limit))
    d["param_unit"] = np.float64(unit)
    d["param_n_families"] = np.int64(len(fam_ids))
    d.update({"ref_" + k: v for k, v in r.items()})
    d["ref_cap_counts"] = caps
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    return len(r["seg_root"]), len(r["hit_root"])


def make_abcd():
    """Fixture for the callers of ray_tracing (optical_table.py:211-422): the reference's own
    calculate_abcd_matrix and calibrate_symmetric_4f(optimize=False) on the LENS-9 asphere with 7 id'd rays."""
    ref = RH.load_reference()
    lens = scenes.asphere_lens9(ref, [43.17, 0, 0])
    rays = scenes.abcd_rays(ref)
    F1, F2 = 46.52, 43.48
    # the body of calibrate_symmetric_4f.simulate (optical_table.py:328-347); the method itself cannot run
    # headless with optimize=False because it ends in table.render(type=None), which raises
    l0 = lens.copy()._Translate(np.array([F1, 0, 0]) - lens.origin)
    l1 = lens.copy()._Translate(np.array([F1 + 2 * F2, 0, 0]) - lens.origin).RotZ(np.pi)
    m0, m1 = ref.Monitor(origin=[0, 0, 0], width=5, height=5), ref.Monitor(origin=[2 * F1 + 2 * F2, 0, 0], width=5, height=5)
    table = ref.OpticalTable()
    table.add_components([l0, l1])
    table.add_monitors([m0, m1])
    table.ray_tracing(rays)
    y, ty = m1.get_yList(), m1.get_tYList()
    Ms = table.calculate_abcd_matrix(m0, m1, rays)
    # the full Nelder-Mead calibration of the example (optimize=True does not render; ~30 s of reference time)
    F_opt = ref.OpticalTable.calibrate_symmetric_4f(lens, scenes.abcd_rays(ref), F10=F1, F20=F2, criterion="M=-I", optimize=True)
    np.savez_compressed(os.path.join(OUT, "abcd_4f.npz"), Ms=np.asarray(Ms), yList=np.asarray(y), tYList=np.asarray(ty),
                        F=np.array([F1, F2]), F_opt=np.asarray(F_opt, dtype=float))
    return np.asarray(Ms)


def make_ripa2_post():
    """Fixture for the analysis of examples/ripa_gen2_2nd_simplified.py:169-203 (Monitor.get_rays / get_ray_id by
    ray id, solve_ray_ray_intersection, Gaussian-beam helpers on the recorded segments), run by the reference."""
    ref = RH.load_reference()
    sc = scenes.ripa2_simplified(ref)
    table = ref.OpticalTable()
    table.add_components(sc.components)
    table.add_monitors(sc.monitors)
    table.ray_tracing(sc.rays)
    post = scenes.ripa2_postprocess(ref, table, sc)
    np.savez_compressed(os.path.join(OUT, "ripa2_post.npz"), **post)
    return post


ANALYTICS_SCENES = ("doublet", "telescope_4f", "callable_material", "ripa2_simplified")


def make_monitor_analytics():
    """Fixture for the monitor analytics (monitor.py:78-253): the reference's OWN Monitor methods -- sorted accessor
    views, get_waist_distance, _get_hist_y / std_histy, get_delta_pos, sum / avg intensity -- evaluated by the
    reference after its own trace, per scene and monitor."""
    ref = RH.load_reference()
    d = {}
    for name in ANALYTICS_SCENES:
        sc = scenes.REGISTRY[name](ref)
        table = ref.OpticalTable()
        table.add_components(sc.components)
        table.add_monitors(sc.monitors)
        table.ray_tracing(sc.rays, perfomance_limit=sc.limit)
        for m, mon in enumerate(table.monitors):
            if mon.ndata < 2:
                continue
            pre = f"{name}__{m}__"
            d[pre + "yList"], d[pre + "zList"] = mon.get_yList(), mon.get_zList()
            d[pre + "tYList"], d[pre + "tZList"] = mon.get_tYList(), mon.tZList
            d[pre + "IList"], d[pre + "tList"] = mon.get_IList(), mon.get_tList()
            d[pre + "waist_distance"] = mon.get_waist_distance()
            counts, edges = mon._get_hist_y()
            d[pre + "hist_counts"], d[pre + "hist_edges"] = counts, edges
            with np.errstate(all="ignore"):
                d[pre + "std_histy"] = np.float64(mon.std_histy)
            dy, dz = mon.get_delta_pos()
            d[pre + "delta_y"], d[pre + "delta_z"] = dy, dz
            d[pre + "sum_intensity"], d[pre + "avg_intensity"] = np.float64(mon.sum_intensity), np.float64(mon.avg_intensity)
    np.savez_compressed(os.path.join(OUT, "monitor_analytics.npz"), **d)
    return sorted({k.rsplit("__", 1)[0] for k in d})


if __name__ == "__main__":
    if sys.argv[1:] == ["analytics"]:
        print(make_monitor_analytics())
        sys.exit(0)
    if sys.argv[1:] == ["abcd"]:
        print(make_abcd()[:2])
        sys.exit(0)
    if sys.argv[1:] == ["ripa2_post"]:
        print({k: v.shape for k, v in make_ripa2_post().items()})
        sys.exit(0)
    for name in (sys.argv[1:] or scenes.REGISTRY):
        nseg, nhit = make(name)
        print(f"{name:20s} segments={nseg} hits={nhit}")
