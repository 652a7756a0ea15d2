"""CPU, build container only: CSV exports (optical_table.py:447-523) against the reference's own, on the
reference's classes and on this package's, with the C oracle standing in for the device."""
import csv
import json

import numpy as np
import pytest

from oracle import ref_harness as RH
from tests import scenes
from tests.test_install_cpu import OracleEngine

pytestmark = pytest.mark.skipif(not RH.reference_available(), reason="/root/reference not present")


def _num(cell):
    """Parse a Mathematica-flavoured cell back into numbers (inverse of base.py:248-251)."""
    if cell == "None" or not isinstance(cell, str):
        return cell
    text = cell.replace("{", "[").replace("}", "]").replace("*10^", "e").replace("I", "j")
    try:
        return np.array(json.loads(text), dtype=float)
    except (ValueError, TypeError):
        pass
    try:
        return complex(text)
    except ValueError:
        return cell  # a name or a class


def _same_rows(want, got, rtol=1e-9):
    assert len(want) == len(got)
    for w, g in zip(want, got):
        assert list(w.keys()) == list(g.keys())
        for k in w:
            a, b = _num(w[k]), _num(g[k])
            if isinstance(a, str) or isinstance(b, str):
                assert a == b, (k, w[k], g[k])
            else:
                assert np.allclose(a, b, rtol=rtol, atol=1e-12), (k, w[k], g[k])


@pytest.mark.parametrize("name", ["doublet", "misc_components", "telescope_4f", "mma_small"])
def test_component_rows_equal_reference(name):
    import optable_b200 as ob

    ref = RH.load_reference()
    a, b = scenes.REGISTRY[name](ref), scenes.REGISTRY[name](ob)
    ta, tb = ref.OpticalTable(), ob.OpticalTable()
    ta.add_components(a.components)
    tb.add_components(b.components)
    _same_rows(ta.gather_components(), tb.gather_components(), rtol=1e-12)
    kw = dict(avoid_flatten_classname=["Doublet", "MMA"], ignore_classname=["Mirror", "Refraction"])
    _same_rows(ta.gather_components(**kw), tb.gather_components(**kw), rtol=1e-12)


@pytest.mark.parametrize("name", ["gaussian_beam", "chromatic", "prism_refl"])
def test_ray_rows_equal_reference(name, tmp_path):
    import optable_b200 as ob
    from optable_b200 import export
    from optable_b200.flatten import FlatScene, pack_rays, trace_cap
    from optable_b200.table import trace_table

    ref = RH.load_reference()
    a, b = scenes.REGISTRY[name](ref), scenes.REGISTRY[name](ob)
    ta, tb = ref.OpticalTable(), ob.OpticalTable()
    for t, sc in ((ta, a), (tb, b)):
        t.add_components(sc.components)
        t.add_monitors(sc.monitors)
    ta.ray_tracing(a.rays, perfomance_limit=a.limit)
    tb.rays.extend(trace_table(tb, list(b.rays), b.limit, engine=OracleEngine()))
    want = ta.gather_rays_csv()
    _same_rows(want, tb.gather_rays_csv())
    # straight from the segment columns, no Ray objects in between
    arrs, fams, unit = pack_rays(b.rays)
    out = OracleEngine().trace_arrays(OracleEngine().upload(FlatScene(tb.components, tb.monitors)), arrs,
                                      max_trace_num=trace_cap(b.limit), unit=unit, n_families=len(fams))
    _same_rows(want, export.segment_rows(out))
    # files: same header, same number of lines, cells parse to the same numbers
    fa, fb = tmp_path / "ref.csv", tmp_path / "own.csv"
    ta.export_rays_csv(str(fa))
    tb.export_rays_csv(str(fb))
    ra, rb = list(csv.reader(open(fa))), list(csv.reader(open(fb)))
    assert ra[0] == rb[0] == list(export.RAY_KEYS) and len(ra) == len(rb)
    _same_rows([dict(zip(ra[0], r)) for r in ra[1:]], [dict(zip(rb[0], r)) for r in rb[1:]])


def test_empty_exports(tmp_path):
    import optable_b200 as ob

    t = ob.OpticalTable()
    t.export_rays_csv(str(tmp_path / "r.csv"))
    t.export_components_csv(str(tmp_path / "c.csv"))
    assert open(tmp_path / "r.csv").read().strip() == "" and open(tmp_path / "c.csv").read().strip() == ""
