"""Text export of traced segments and of the scene description (reference: optable/optical_table.py:447-523,
optable/optical_component.py:386-426, string helpers optable/base.py:248-259).

The reference writes CSV files whose cells are Mathematica-flavoured strings (`{a, b}` lists, `*10^` exponents,
`I` for the imaginary unit). `segment_rows` produces the same cells straight from the device's segment columns,
so an export does not need `Ray` objects; `ray_rows` is the object-level twin used by
`OpticalTable.gather_rays_csv`.
"""
from __future__ import annotations

import csv

import numpy as np

from . import _abi as A
from .pose import rotation_from_x

RAY_KEYS = ("origin", "transform_matrix", "intensity", "length", "qo", "n")
COMPONENT_KEYS = ("name", "class", "origin", "transform_matrix", "radius", "width", "height", "focal_length")


def mathematical_str(text: str) -> str:
    """'[1e-05, 2j]' -> '{1*10^-05, 2I}' (base.py:248-251)."""
    if text == "None":
        return "None"
    return text.translate(_MATH_TABLE)


_MATH_TABLE = {ord("["): "{", ord("]"): "}", ord("e"): "*10^", ord("j"): "I"}


def _truthy_or_none(obj, name):
    """The attribute if present and truthy, else the string 'None' (base.py:254-259: zero also prints None)."""
    value = getattr(obj, name, None)
    return value if value else "None"


def _cell_vec(v) -> str:
    return mathematical_str(str(np.asarray(v).tolist()))


def ray_rows(rays) -> list:
    """One dict per ray object, keys RAY_KEYS (optical_table.py:448-469)."""
    rows = []
    for r in rays:
        rows.append({
            "origin": _cell_vec(r.origin),
            "transform_matrix": _cell_vec(r.transform_matrix),
            "intensity": _truthy_or_none(r, "intensity"),
            "length": _truthy_or_none(r, "length"),
            "qo": mathematical_str(str(_truthy_or_none(r, "qo"))),
            "n": mathematical_str(str(_truthy_or_none(r, "n"))),
        })
    return rows


def segment_rows(seg: dict) -> list:
    """The same rows from segment columns (`Engine.trace_arrays` output, (root, pop) order)."""
    n = len(seg["seg_root"])
    O = np.stack((seg["seg_ox"], seg["seg_oy"], seg["seg_oz"]), 1).tolist()
    D = np.stack((seg["seg_dx"], seg["seg_dy"], seg["seg_dz"]), 1)
    inten, length, nmed = seg["seg_intensity"].tolist(), seg["seg_length"].tolist(), seg["seg_n"].tolist()
    hasq = ((seg["seg_flags"] & A.RF_HASQ) != 0).tolist()
    q = (seg["seg_q_re"] + 1j * seg["seg_q_im"]).tolist()
    rows = []
    for k in range(n):
        finite = length[k] != float("inf")
        rows.append({
            "origin": mathematical_str(str(O[k])),
            "transform_matrix": _cell_vec(rotation_from_x(D[k])),
            "intensity": inten[k] if inten[k] else "None",
            "length": length[k] if (finite and length[k]) else "None",
            "qo": mathematical_str(str(q[k] if (hasq[k] and q[k]) else "None")),
            "n": mathematical_str(str(nmed[k] if nmed[k] else "None")),
        })
    return rows


def component_rows(component, avoid_flatten_classname=(), ignore_classname=()) -> list:
    """Depth-first description of a component tree (optical_component.py:386-426): the node itself unless its
    class is ignored, then its children unless its class is listed as not to be expanded."""
    cls = type(component).__name__
    rows = []
    if cls not in ignore_classname:
        rows.append({
            "name": _truthy_or_none(component, "name"),
            "class": cls,
            "origin": _cell_vec(component.origin),
            "transform_matrix": _cell_vec(component.transform_matrix),
            "radius": _truthy_or_none(component, "radius"),
            "width": _truthy_or_none(component, "width"),
            "height": _truthy_or_none(component, "height"),
            "focal_length": _truthy_or_none(component, "focal_length"),
        })
    if cls not in avoid_flatten_classname:
        for child in getattr(component, "components", ()):
            rows.extend(component_rows(child, avoid_flatten_classname, ignore_classname))
    return rows


def write_csv(filename: str, rows: list) -> None:
    """Header from the first row's keys, one line per row (optical_table.py:484-497); empty input -> empty header."""
    with open(filename, "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(rows[0].keys() if rows else [])
        for row in rows:
            w.writerow(row.values())
