"""GPU: the CUDA bounce loop against (a) the real reference's results in tests/golden/ and (b) the C oracle."""
import numpy as np
import pytest

from tests import golden_io, parity

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    from optable_b200.backend import Engine

    return Engine.get(0)


_q_rtol = parity.q_rtol_for


def _gpu(engine, flat, rays, params, **kw):
    from oracle import ref_harness as RH

    scene = engine.upload(flat)
    out = engine.trace_arrays(scene, rays, **params, **kw)
    return out, RH.arrays_from_result(out)


@pytest.mark.parametrize("name", golden_io.names())
def test_cuda_matches_reference_golden(engine, name):
    flat, rays, params, ref = golden_io.load(name)
    out, got = _gpu(engine, flat, rays, params)
    parity.compare(ref, got, q_rtol=_q_rtol(flat), label=name)
    if flat.n_capslots:
        np.testing.assert_array_equal(out["cap_counts"], ref["cap_counts"])
    assert int(out["counters"][1]) == int((ref["seg_leaf"] >= 0).sum())


@pytest.mark.parametrize("name", ["gaussian_beam", "cavity_aligned", "doublet", "telescope_4f", "misc_components"])
@pytest.mark.parametrize("chain_len", [1, 3])
def test_wavefront_scheduling_does_not_change_results(engine, name, chain_len):
    """chain_len only moves work between the in-register loop and the wavefront queue."""
    flat, rays, params, ref = golden_io.load(name)
    _, got = _gpu(engine, flat, rays, params, chain_len=chain_len)
    parity.compare(ref, got, q_rtol=_q_rtol(flat), label=f"{name}/chain{chain_len}")


@pytest.mark.parametrize("block", range(6))
def test_fuzz_scenes_cuda_equals_oracle(engine, block):
    """300 random scenes over the whole component zoo (tests/scenes.fuzz, 64 rays each, built with this package's
    classes), CUDA path through the C ABI against the C oracle: indices and pop counts exact, fields to 1e-9."""
    import optable_b200 as ob
    from optable_b200.flatten import FlatScene, pack_rays, trace_cap
    from oracle import oracle as O
    from oracle import ref_harness as RH
    from tests import scenes

    pops = 0
    for seed in range(1000 + 50 * block, 1050 + 50 * block):
        sc = scenes.fuzz(ob, seed, n_rays=64)
        flat = FlatScene(sc.components, sc.monitors)
        arrs, fam_ids, unit = pack_rays(sc.rays)
        params = dict(max_trace_num=trace_cap(sc.limit), unit=unit, n_families=len(fam_ids))
        want = RH.arrays_from_result(O.trace(flat, arrs, **params))
        _, got = _gpu(engine, flat, arrs, params)
        parity.compare(want, got, q_rtol=_q_rtol(flat), label=f"fuzz seed {seed}")
        pops += len(want["seg_root"])
    assert pops > 5000
