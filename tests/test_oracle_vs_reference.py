"""CPU, build container only: the C restatement against the live Python reference on every fixture scene."""
import pytest

from oracle import oracle as O
from oracle import ref_harness as RH
from optable_b200.flatten import FlatScene, pack_rays, trace_cap
from tests import parity, scenes

pytestmark = pytest.mark.skipif(not RH.reference_available(), reason="/root/reference not present")


@pytest.mark.parametrize("name", list(scenes.REGISTRY))
def test_oracle_vs_live_reference(name, capsys):
    ref = RH.load_reference()
    sc = scenes.REGISTRY[name](ref)
    flat = FlatScene(sc.components, sc.monitors)
    arrs, fam_ids, unit = pack_rays(sc.rays)
    want = RH.run_reference(sc)
    got = O.trace(flat, arrs, max_trace_num=trace_cap(sc.limit), unit=unit, n_families=len(fam_ids))
    parity.compare(want, RH.arrays_from_result(got), label=name)
