"""Run the REAL reference (tim4431/optable, Python) in the build container.

TEST INFRASTRUCTURE ONLY. /root/reference does not exist on the GPU box; what travels there is (a) the
fixtures this harness produced (tests/golden/, made by oracle/make_golden.py) and (b) the UNMODIFIED reference
package as `pip install --target baseline/_ref` put it (git-ignored, done by __graft_entry__.build()),
which the `-m gpu` drop-in tests and the `kind: "reference"` leg of bench.py import from there. The reference is
imported unmodified; matplotlib (not installed here, never called on the hot path) is replaced by empty stub
modules.
"""
from __future__ import annotations

import math
import os
import sys
import types

import numpy as np

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INSTALLED_ROOT = os.path.join(_REPO, "baseline", "_ref")  # pip --target copy of the reference (travels with gpurun)


def _pick_root():
    for cand in (os.environ.get("OPTABLE_REFERENCE_ROOT"), "/root/reference", INSTALLED_ROOT):
        if cand and os.path.isdir(os.path.join(cand, "optable")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _pick_root()
_ref = None


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "optable"))


def install_reference(force=False) -> str:
    """`pip install --no-index --no-deps --target baseline/_ref` of the reference tree (from a /tmp copy: the
    tree itself is read-only). Returns the target directory, or "" when /root/reference is absent (GPU box: the
    copy made in the build container travelled with the snapshot)."""
    import shutil
    import subprocess
    import tempfile

    if os.path.isdir(os.path.join(INSTALLED_ROOT, "optable")) and not force:
        return INSTALLED_ROOT
    src = "/root/reference"
    if not os.path.isdir(os.path.join(src, "optable")):
        return ""
    tmp = tempfile.mkdtemp(prefix="optable_src_")
    try:
        work = os.path.join(tmp, "src")
        shutil.copytree(src, work, ignore=shutil.ignore_patterns(".git", "docs"))
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--quiet",
               "--find-links", "/opt/wheelhouse", "--upgrade", "--target", INSTALLED_ROOT, work]
        subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return INSTALLED_ROOT


def load_reference():
    """Import the reference package under the private name `optable` and return the module."""
    global _ref
    if _ref is not None:
        return _ref
    if not reference_available():
        raise RuntimeError("reference tree not present")
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.lines", "matplotlib.gridspec", "mpl_toolkits",
                 "mpl_toolkits.mplot3d", "mpl_toolkits.mplot3d.art3d"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["mpl_toolkits.mplot3d.art3d"].Poly3DCollection = object
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import optable  # noqa: the reference

    _ref = optable
    return _ref


def _leaves(components, out):
    for c in components:
        if hasattr(c, "components"):
            _leaves(c.components, out)
        else:
            out.append(c)
    return out


def run_reference(scene):
    """Trace `scene` (tests/scenes.Scene built with the reference namespace) with the reference's own
    OpticalTable.ray_tracing, one initial ray at a time, and return numpy arrays in the layout of
    include/optb.h results: segments in (root, pop) order, hits in (root, monitor, pop) order.
    The winning leaf of every pop is recovered by wrapping each leaf's bound `interact` (instance
    attribute; the reference sources are untouched)."""
    ref = load_reference()
    table = ref.OpticalTable()
    table.add_components(scene.components)
    table.add_monitors(scene.monitors)
    # dense leaf numbering identical to FlatScene (PointObj / Point surfaces are skipped there)
    leaves = [c for c in _leaves(table.components, []) if type(c.surface).__name__ != "Point"]
    log = []

    def wrap(idx, comp):
        orig = comp.interact

        def interact(ray):
            t, rays = orig(ray)
            if t is not None:
                log.append((id(ray), idx, float(t)))
            return t, rays

        comp.interact = interact

    for k, c in enumerate(leaves):
        wrap(k, c)

    seg = {k: [] for k in ("o", "d", "length", "alive", "I", "wl", "q", "hasq", "pl", "n", "root", "pop", "leaf")}
    hits = {k: [] for k in ("mon", "root", "pop", "P", "I", "t", "d", "q")}
    for root, ray in enumerate(scene.rays):
        n0 = len(table.rays)
        m0 = [len(m._data_raw) for m in table.monitors]
        del log[:]
        table.ray_tracing([ray], perfomance_limit=scene.limit)
        new = table.rays[n0:]
        # winners: consecutive log entries with the same ray object belong to one pop; smallest t, first wins
        winners, cur, best = [], None, None
        for rid, idx, t in log:
            if rid != cur:
                if best is not None:
                    winners.append(best)
                cur, best = rid, (idx, t)
            elif t < best[1]:
                best = (idx, t)
        if best is not None:
            winners.append(best)
        wi = 0
        index_of = {}
        for pop, r in enumerate(new):
            index_of[id(r)] = pop
            seg["o"].append(np.array(r.origin, float)); seg["d"].append(np.array(r.direction, float))
            seg["length"].append(math.inf if r.length is None else float(r.length))
            seg["alive"].append(bool(r.alive)); seg["I"].append(float(r.intensity)); seg["wl"].append(float(r.wavelength))
            seg["hasq"].append(r.qo is not None); seg["q"].append(complex(r.qo) if r.qo is not None else 0j)
            seg["pl"].append(float(r._pathlength)); seg["n"].append(float(r.n))
            seg["root"].append(root); seg["pop"].append(pop)
            if not r.alive and ray.alive:
                idx, t = winners[wi]; wi += 1
                assert t == r.length, (t, r.length)
                seg["leaf"].append(idx)
            else:
                seg["leaf"].append(-1)
        assert wi == len(winners), (wi, len(winners))
        for mi, m in enumerate(table.monitors):
            for (P, I, t, r) in m._data_raw[m0[mi]:]:
                hits["mon"].append(mi); hits["root"].append(root); hits["pop"].append(index_of[id(r)])
                hits["P"].append(np.array(P, float)); hits["I"].append(float(I)); hits["t"].append(float(t))
                hits["d"].append(np.array(r.direction, float)); hits["q"].append(complex(r.qo) if r.qo is not None else 0j)
    for c in leaves:
        del c.interact
    out = {
        "seg_o": np.array(seg["o"], float).reshape(-1, 3), "seg_d": np.array(seg["d"], float).reshape(-1, 3),
        "seg_length": np.array(seg["length"], float), "seg_alive": np.array(seg["alive"], bool),
        "seg_intensity": np.array(seg["I"], float), "seg_wavelength": np.array(seg["wl"], float),
        "seg_q": np.array(seg["q"], complex), "seg_hasq": np.array(seg["hasq"], bool),
        "seg_pathlength": np.array(seg["pl"], float), "seg_n": np.array(seg["n"], float),
        "seg_root": np.array(seg["root"], np.int64), "seg_pop": np.array(seg["pop"], np.int64),
        "seg_leaf": np.array(seg["leaf"], np.int64),
        "hit_monitor": np.array(hits["mon"], np.int64), "hit_root": np.array(hits["root"], np.int64),
        "hit_pop": np.array(hits["pop"], np.int64), "hit_P": np.array(hits["P"], float).reshape(-1, 3),
        "hit_intensity": np.array(hits["I"], float), "hit_t": np.array(hits["t"], float),
        "hit_d": np.array(hits["d"], float).reshape(-1, 3), "hit_q": np.array(hits["q"], complex),
    }
    # interact counts per capped leaf and family id, after the trace
    out["_leaves"] = leaves
    return out


def arrays_from_result(arrs):
    """optb result arrays (oracle or CUDA, already in reference order) -> the layout of run_reference."""
    n = len(arrs["seg_ox"])
    flags = arrs["seg_flags"]
    out = {
        "seg_o": np.stack([arrs["seg_ox"], arrs["seg_oy"], arrs["seg_oz"]], 1) if n else np.zeros((0, 3)),
        "seg_d": np.stack([arrs["seg_dx"], arrs["seg_dy"], arrs["seg_dz"]], 1) if n else np.zeros((0, 3)),
        "seg_length": arrs["seg_length"], "seg_alive": (flags & 1) != 0,
        "seg_intensity": arrs["seg_intensity"], "seg_wavelength": arrs["seg_wavelength"],
        "seg_q": arrs["seg_q_re"] + 1j * arrs["seg_q_im"], "seg_hasq": (flags & 2) != 0,
        "seg_pathlength": arrs["seg_pathlength"], "seg_n": arrs["seg_n"],
        "seg_root": arrs["seg_root"].astype(np.int64), "seg_pop": arrs["seg_pop"].astype(np.int64),
        "seg_leaf": arrs["seg_leaf"].astype(np.int64),
    }
    m = len(arrs["hit_px"])
    out.update({
        "hit_monitor": arrs["hit_monitor"].astype(np.int64), "hit_root": arrs["hit_root"].astype(np.int64),
        "hit_pop": arrs["hit_pop"].astype(np.int64),
        "hit_P": np.stack([arrs["hit_px"], arrs["hit_py"], arrs["hit_pz"]], 1) if m else np.zeros((0, 3)),
        "hit_intensity": arrs["hit_intensity"], "hit_t": arrs["hit_t"],
        "hit_d": np.stack([arrs["hit_dx"], arrs["hit_dy"], arrs["hit_dz"]], 1) if m else np.zeros((0, 3)),
        "hit_q": arrs["hit_q_re"] + 1j * arrs["hit_q_im"],
    })
    return out
