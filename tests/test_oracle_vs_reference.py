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
    q_rtol = parity.q_rtol_for(flat) if name.startswith("fuzz_") else parity.RTOL
    parity.compare(want, RH.arrays_from_result(got), q_rtol=q_rtol, label=name)


@pytest.mark.parametrize("block", range(6))
def test_fuzz_scenes_oracle_equals_reference(block):
    """Random scenes over the whole component zoo (tests/scenes.fuzz): 60 scenes here; 400 were run when the
    generator was written (72,274 segments, no index, pop-count or tolerance difference)."""
    ref = RH.load_reference()
    for seed in range(100 + 10 * block, 110 + 10 * block):
        sc = scenes.fuzz(ref, seed)
        flat = FlatScene(sc.components, sc.monitors)
        arrs, fam_ids, unit = pack_rays(sc.rays)
        want = RH.run_reference(sc)
        got = O.trace(flat, arrs, max_trace_num=trace_cap(sc.limit), unit=unit, n_families=len(fam_ids))
        parity.compare(want, RH.arrays_from_result(got), q_rtol=parity.q_rtol_for(flat), label=f"fuzz seed {seed}")
