"""GPU: the literal drop-in. `optable_b200.install(reference)` swaps `OpticalTable.ray_tracing` of the UNMODIFIED
reference package (the `pip install --target baseline/_ref` copy that travels with the snapshot) for the
CUDA engine; the same scene, built twice from the reference's OWN classes, is traced once by the reference's
original method on the host and once through liboptb.so, and compared object by object:
`table.rays`, `Monitor._data_raw`, `_interact_count`, the monitors' accessors, `calculate_abcd_matrix`
(/root/reference/optable/optical_table.py:57-147, 211-297)."""
import numpy as np
import pytest

from oracle import ref_harness as RH
from tests import parity, scenes
from tests.install_check import check_install

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not RH.reference_available(), reason="baseline/_ref (pip --target copy of the reference) not present")]


def _engine():
    from optable_b200.backend import Engine

    return Engine.get(0)


def _q_rtol(name):
    """q after an ASphere surface: the reference's finite-difference curvature amplifies rounding."""
    from optable_b200.flatten import FlatScene
    import optable_b200 as ob

    sc = scenes.REGISTRY[name](ob)
    return parity.q_rtol_for(sc.flat())


@pytest.mark.parametrize("name", sorted(scenes.REGISTRY))
def test_cuda_engine_under_reference_classes(name):
    n_seg, n_rows = check_install(name, _engine(), q_rtol=_q_rtol(name))
    assert n_seg > 0


def test_reference_abcd_matrix_on_cuda_backend():
    """The reference's OWN `calculate_abcd_matrix` (three `ray_tracing` calls, monitor accessors sorted by ray id,
    optical_table.py:211-297) running on top of the swapped back end equals itself on the original back end."""
    import optable_b200

    ref = RH.load_reference()

    def build():
        F1 = F2 = 43.17
        lens = scenes.asphere_lens9(ref, [F1, 0, 0])
        l1 = scenes.asphere_lens9(ref, [F1 + 2 * F2, 0, 0]).RotZ(np.pi)
        m0, m1 = ref.Monitor([0, 0, 0], width=5, height=5), ref.Monitor([2 * F1 + 2 * F2, 0, 0], width=5, height=5)
        t = ref.OpticalTable()
        t.add_components([lens, l1])
        t.add_monitors([m0, m1])
        return t, m0, m1

    ta, a0, a1 = build()
    want = ta.calculate_abcd_matrix(a0, a1, scenes.abcd_rays(ref))
    tb, b0, b1 = build()
    original = optable_b200.install(ref, engine=_engine())
    try:
        got = tb.calculate_abcd_matrix(b0, b1, scenes.abcd_rays(ref))
    finally:
        ref.OpticalTable.ray_tracing = original
    # finite differences of traced positions over 1e-5: 1e-9 trace parity -> ~1e-4 absolute on the entries
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=2e-4)
    assert len(ta.rays) == len(tb.rays) and len(a1._data_raw) == len(b1._data_raw)
    np.testing.assert_allclose(b1.get_yList(sort="ID"), a1.get_yList(sort="ID"), rtol=1e-9, atol=1e-12)


def test_native_library_is_what_ran():
    """The swapped method must run liboptb.so (no CPU path exists): the engine is the ctypes-bound CUDA context."""
    from optable_b200 import backend

    eng = _engine()
    assert isinstance(eng, backend.Engine) and backend.lib().optb_abi_version() > 0
    import ctypes

    maps = open("/proc/self/maps").read()
    assert "liboptb.so" in maps
