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


# whole random paths (up to 60 pops through trapped, splitting geometry): every field except q to 1e-7 (observed
# worst over 400 scenes: 3.7e-8 on one length of the extended zoo, 8e-10 over the basic zoo, 2e-11 on positions /
# directions); q in scenes whose curvature comes from the reference's finite-difference stencil: FUZZ_Q_FD
FUZZ_PATH_RTOL = 1e-7
FUZZ_Q_FD = 2e-5   # (the finite-difference noise of SURVEY A.11 grows with |q| / ROC; worst seen over 400 scenes: 1.2e-5)
DEVICE_FLAGS = {"rays": 0, "flagged": 0}  # in-kernel A.9 flags over the fuzz scenes (reported by the last fuzz test)


def _fuzz_block(engine, seeds, **fuzz_kw):
    """Whole paths (ties flagged, loose fields) + restarted single interactions (strict) for a range of fuzz seeds;
    returns (pops, restarted rays, flagged roots, rays compared)."""
    import optable_b200 as ob
    from optable_b200.flatten import FlatScene, pack_rays, trace_cap
    from oracle import oracle as O
    from oracle import ref_harness as RH
    from tests import scenes

    pops = flagged = rays = restarted = 0
    for seed in seeds:
        sc = scenes.fuzz(ob, seed, n_rays=64, **fuzz_kw)
        flat = sc.flat()
        arrs, fam_ids, unit = pack_rays(sc.rays)
        params = dict(max_trace_num=trace_cap(sc.limit), unit=unit, n_families=len(fam_ids))
        raw = O.trace(flat, arrs, **params)
        want = RH.arrays_from_result(raw)
        _, got = _gpu(engine, flat, arrs, params)
        _, ties = parity.compare_flagging_ties(flat, want, got, rtol=FUZZ_PATH_RTOL, q_rtol=FUZZ_Q_FD if _q_rtol(flat) > parity.RTOL else FUZZ_PATH_RTOL,
                                               label=f"fuzz seed {seed}")
        pops += len(want["seg_root"])
        flagged += len(ties)
        rays += len(sc.rays)
        if not flat.n_capslots:
            # the device's own ambiguity mask (params.flag_ambiguity): every root whose two traces part
            # ways at an equal-distance tie must carry a bit, and flagged rays stay a small minority
            scene = engine.upload(flat)
            fl = engine.trace_arrays(scene, arrs, flag_ambiguity=True, **params)["root_flags"]
            scene.close()
            assert all(fl[r] != 0 for r in ties), (seed, ties, [int(fl[r]) for r in ties])
            DEVICE_FLAGS["rays"] += len(fl)
            DEVICE_FLAGS["flagged"] += int((fl != 0).sum())
        batch = parity.restart_batch(raw, np.nonzero(np.isinf(arrs["length"]))[0])
        p1 = dict(max_trace_num=3, unit=unit, n_families=len(batch["ox"]))
        want1 = RH.arrays_from_result(O.trace(flat, batch, **p1))
        _, got1 = _gpu(engine, flat, batch, p1)
        # (q behind a finite-difference asphere curvature: the 1e-6 carve-out of the fixtures grows to 1e-5 once
        # random upstream optics have made |q| hundreds of times the radius of curvature)
        q1 = FUZZ_Q_FD if _q_rtol(flat) > parity.RTOL else parity.RTOL
        _, ties1 = parity.compare_flagging_ties(flat, want1, got1, q_rtol=q1, label=f"fuzz seed {seed} (restarted)")
        # the same single interactions with the reference's own root iteration (params.reference_roots: brentq_dev on
        # the reference's brackets): the device lands on brentq's last iterate instead of on the true root, and the
        # bar on every geometric field drops from 1e-9 to 1e-10 (observed 7e-12); decisions are the default mode's
        _, got1b = _gpu(engine, flat, batch, p1, reference_roots=True)
        if not ties1:
            parity.compare(want1, got1b, rtol=1e-10, q_rtol=q1, label=f"fuzz seed {seed} (restarted, reference roots)")
            for k in ("seg_leaf", "seg_pop", "hit_monitor"):
                assert np.array_equal(got1[k], got1b[k]), (seed, k)
        restarted += len(batch["ox"])
        flagged += len(ties1)
        rays += len(batch["ox"])
    return pops, restarted, flagged, rays


@pytest.mark.parametrize("block", range(6))
def test_fuzz_scenes_cuda_equals_oracle(engine, block):
    """300 random scenes over the component zoo (tests/scenes.fuzz, 64 rays each, built with this package's
    classes), CUDA path through the C ABI against the C oracle.
    (a) Whole paths (up to 60 pops, splitting): winning leaf, pop numbering and segment counts exact (equal-distance
        ties flagged, see parity.compare_flagging_ties); fields to 1e-6 (q 1e-5), because random scenes trap rays between
        curved faces where the reference's own root tolerance (brentq xtol = 2e-12 absolute) is amplified per bounce.
    (b) Single interactions: every popped ray of (a) restarted on both sides from identical inputs and traced for
        three pops: the strict 1e-9 bar on every field."""
    pops, restarted, flagged, rays = _fuzz_block(engine, range(1000 + 50 * block, 1050 + 50 * block))
    # equal-distance ties between coplanar overlapping apertures: a few per 1e4 rays
    assert pops > 5000 and restarted > 4000 and flagged <= 2e-3 * rays, (flagged, rays)


@pytest.mark.parametrize("block", range(2))
def test_fuzz_scenes_whole_zoo_cuda_equals_oracle(engine, block):
    """The same two comparisons over the extended zoo (tests/scenes.fuzz(extended=True): + MMA, MirrorPair,
    MirrorPrism, TriangularPrism with its built-in caps, DovePrism polygons, the exact-spherical asphere, bare
    refractive faces)."""
    pops, restarted, flagged, rays = _fuzz_block(engine, range(50000 + 50 * block, 50050 + 50 * block), extended=True)
    assert pops > 5000 and restarted > 4000 and flagged <= 2e-3 * rays, (flagged, rays)
    if block == 1:  # random scenes are full of edges and overlapping apertures; even so the mask stays a minority
        print("device A.9 flags over the fuzz scenes:", DEVICE_FLAGS)
        assert DEVICE_FLAGS["rays"] > 10000 and DEVICE_FLAGS["flagged"] <= 0.2 * DEVICE_FLAGS["rays"], DEVICE_FLAGS


def test_fuzz_scenes_with_binding_interact_caps(engine):
    """Random scenes with binding max_interact_count and ray families sharing an id (tests/scenes.fuzz(caps=True)):
    the family-serial kernel against the oracle, interact-count tables included."""
    import optable_b200 as ob
    from optable_b200.flatten import FlatScene, pack_rays, trace_cap
    from oracle import oracle as O
    from oracle import ref_harness as RH
    from tests import scenes

    reached = flagged = 0
    for seed in range(300, 340):
        sc = scenes.fuzz(ob, seed, caps=True)
        flat = sc.flat()
        arrs, fam_ids, unit = pack_rays(sc.rays)
        params = dict(max_trace_num=trace_cap(sc.limit), unit=unit, n_families=len(fam_ids))
        raw = O.trace(flat, arrs, **params)
        out, got = _gpu(engine, flat, arrs, params)
        _, ties = parity.compare_flagging_ties(flat, RH.arrays_from_result(raw), got, rtol=FUZZ_PATH_RTOL,
                                               q_rtol=FUZZ_Q_FD if _q_rtol(flat) > parity.RTOL else FUZZ_PATH_RTOL, label=f"caps fuzz seed {seed}")
        flagged += len(ties)
        if flat.n_capslots and not ties:
            np.testing.assert_array_equal(out["cap_counts"], raw["cap_counts"])
            capmax = flat.node_f[flat.node_i[:, 6] >= 0, 39]
            reached += int((raw["cap_counts"].max(axis=1) >= capmax).any())
    assert reached > 20 and flagged <= 3


def test_random_nested_composite_apertures(engine):
    """120 random composites of circles, rectangles and polygons nested up to four operator levels
    (tests/scenes.random_csg; the CPU suite pins the oracle to the live reference on the first 40): the device's bit-stack
    evaluation of the flattener's postfix programs against the oracle, strict bars (one interaction deep, nothing to
    amplify), also with the ambiguity mask on (its own copy of the aperture test)."""
    import optable_b200 as ob
    from optable_b200 import _abi as A
    from optable_b200.flatten import pack_rays, trace_cap
    from oracle import oracle as O
    from oracle import ref_harness as RH
    from tests import scenes

    nested = rows = 0
    for seed in range(120):
        sc = scenes.random_csg(ob, seed)
        flat = sc.flat()
        nested += int(((flat.node_i[:, A.NI_GEOM] == A.G_CSG) & (flat.node_f[:, A.NF_P] == 2.0)).sum())
        arrs, fam_ids, unit = pack_rays(sc.rays)
        params = dict(max_trace_num=trace_cap(sc.limit), unit=unit, n_families=len(fam_ids))
        want = RH.arrays_from_result(O.trace(flat, arrs, **params))
        _, got = _gpu(engine, flat, arrs, params)
        _, ties = parity.compare_flagging_ties(flat, want, got, rtol=parity.RTOL, q_rtol=parity.RTOL, label=f"random_csg seed {seed}")
        assert not ties, (seed, ties)
        if seed % 4 == 0:
            _, got_f = _gpu(engine, flat, arrs, params, flag_ambiguity=True)
            parity.compare(want, got_f, label=f"random_csg seed {seed} (flag variant)")
        rows += len(want["hit_root"])
    assert nested >= 90 and rows > 1000
