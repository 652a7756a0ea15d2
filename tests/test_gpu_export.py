"""GPU: the segment stream -> CSV exports of the reference (optical_table.py:447-523) fed by the
CUDA engine: cell for cell the numbers the unmodified reference (baseline/_ref) writes for the same scene, both
through Ray objects (`OpticalTable.export_rays_csv`) and straight from the device's segment columns
(`export.segment_rows`), plus the component table."""
import csv

import numpy as np
import pytest

from oracle import ref_harness as RH
from tests import scenes
from tests.test_export_cpu import _same_rows

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not RH.reference_available(), reason="baseline/_ref (pip --target copy of the reference) not present")]


@pytest.mark.parametrize("name", ["gaussian_beam", "chromatic", "prism_refl", "doublet", "callable_material"])
def test_ray_csv_rows_from_the_device_equal_reference(name, tmp_path):
    import optable_b200 as ob
    from optable_b200 import export
    from optable_b200.backend import Engine
    from optable_b200.flatten import pack_rays, trace_cap

    ref = RH.load_reference()
    a, b = scenes.REGISTRY[name](ref), scenes.REGISTRY[name](ob)
    ta, tb = ref.OpticalTable(), ob.OpticalTable()
    for t, sc in ((ta, a), (tb, b)):
        t.add_components(sc.components)
        t.add_monitors(sc.monitors)
    ta.ray_tracing(a.rays, perfomance_limit=a.limit)      # the reference, on the host
    tb.ray_tracing(b.rays, perfomance_limit=b.limit)      # liboptb.so
    want = ta.gather_rays_csv()
    _same_rows(want, tb.gather_rays_csv())
    # straight from the device's segment columns, no Ray objects in between
    engine = Engine.get(0)
    arrs, fams, unit = pack_rays(b.rays)
    scene = engine.upload(b.flat())
    out = engine.trace_arrays(scene, arrs, max_trace_num=trace_cap(b.limit), unit=unit, n_families=len(fams))
    scene.close()
    _same_rows(want, export.segment_rows(out))
    fa, fb = tmp_path / "ref.csv", tmp_path / "own.csv"
    ta.export_rays_csv(str(fa))
    tb.export_rays_csv(str(fb))
    ra, rb = list(csv.reader(open(fa))), list(csv.reader(open(fb)))
    assert ra[0] == rb[0] == list(export.RAY_KEYS) and len(ra) == len(rb) > 1
    _same_rows([dict(zip(ra[0], r)) for r in ra[1:]], [dict(zip(rb[0], r)) for r in rb[1:]])
    _same_rows(ta.gather_components(), tb.gather_components(), rtol=1e-12)
