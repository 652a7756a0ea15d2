"""Load the committed golden fixtures (tests/golden/*.npz, produced by oracle/make_golden.py)."""
import glob
import os

import numpy as np

from optable_b200.flatten import FlatScene

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


NON_SCENE = {"abcd_4f", "ripa2_post", "monitor_analytics"}  # fixtures of callers, not of a single trace


def names():
    found = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    return [n for n in found if n not in NON_SCENE]


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    flat = FlatScene.from_arrays({k[len("scene_"):]: z[k] for k in z.files if k.startswith("scene_")})
    rays = {k[len("ray_"):]: np.ascontiguousarray(z[k]) for k in z.files if k.startswith("ray_")}
    ref = {k[len("ref_"):]: z[k] for k in z.files if k.startswith("ref_")}
    params = dict(max_trace_num=int(z["param_max_trace_num"]), unit=float(z["param_unit"]),
                  n_families=int(z["param_n_families"]))
    return flat, rays, params, ref
