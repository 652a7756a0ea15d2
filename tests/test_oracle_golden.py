"""CPU: the C restatement (oracle/) reproduces the real reference's results stored in tests/golden/."""
import numpy as np
import pytest

from oracle import oracle as O
from oracle import ref_harness as RH
from tests import golden_io, parity


@pytest.mark.parametrize("name", golden_io.names())
def test_oracle_matches_golden(name):
    flat, rays, params, ref = golden_io.load(name)
    got = O.trace(flat, rays, **params)
    q_rtol = parity.q_rtol_for(flat) if name.startswith("fuzz_") else parity.RTOL  # (rotated aspheres only occur there)
    errs = parity.compare(ref, RH.arrays_from_result(got), q_rtol=q_rtol, label=name)
    assert all(v <= (q_rtol if k == "seg_q" else parity.RTOL) for k, v in errs.items())
    if flat.n_capslots:
        np.testing.assert_array_equal(got["cap_counts"], ref["cap_counts"])
    n_inter = int(np.isfinite(ref["seg_length"][ref["seg_leaf"] >= 0]).sum())
    assert got["counters"][1] == n_inter  # OPTB_C_INTERACTIONS


def test_known_answers_survey_appendix_b():
    """SURVEY.md Appendix B.1/B.2 vectors captured from the reference."""
    flat, rays, params, ref = golden_io.load("gaussian_beam")
    got = RH.arrays_from_result(O.trace(flat, rays, **params))
    assert len(got["seg_root"]) == 13
    np.testing.assert_allclose(got["seg_d"][1], [-0.5, -0.8660254037844386, 0.0], rtol=0, atol=1e-15)
    assert got["seg_q"][3] == pytest.approx(-9.99999996755554 + 0.0004027682863082486j, rel=1e-12)
    assert got["seg_pathlength"][3] == 0.0  # thin lens leaves the path length untouched
    assert got["seg_n"][9] == 2.0 and got["seg_length"][9] == 5.0
    assert got["seg_q"][10] == pytest.approx(7.5 + 0.00040276828892176836j, rel=1e-12)
    flat, rays, params, ref = golden_io.load("doublet")
    got = RH.arrays_from_result(O.trace(flat, rays, **params))
    lens = got["seg_length"][got["seg_root"] == 0]
    np.testing.assert_allclose(lens[:3], [30.5327376242269, 1.047370294377616, 0.717031170966231], rtol=1e-12)
    np.testing.assert_allclose(got["seg_d"][3], [0.9972279441376599, -0.0665517989597667, 0.03327589947988335], rtol=1e-12)


def test_oracle_multithreaded_equals_single():
    flat, rays, params, _ = golden_io.load("misc_components")
    a = O.trace(flat, rays, nthreads=1, **params)
    b = O.trace(flat, rays, nthreads=4, **params)
    for k in ("seg_ox", "seg_length", "seg_root", "seg_pop", "hit_px", "hit_monitor"):
        np.testing.assert_array_equal(a[k], b[k])
