"""Canonical digests of a LevelDump (see tests/golden/make_golden.py)."""
import hashlib

import numpy as np


def _h(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def input_digest(dist, mat, blend):
    return _h(dist, mat, blend)


def level_digests(lv):
    """exact = everything that must be bit-exact; normals separately (contract tolerance 1e-5, 0 ULP observed)."""
    r = lv.rows
    return {
        "counts": [int(len(r)), int(len(lv.verts)), int(len(lv.idx)), int(len(lv.tverts)), int(len(lv.tidx))],
        "exact": _h(r["id"], r["min"], r["max"], r["nv"], r["ni"], r["tnv"], r["tni"],
                    lv.verts["pos"], lv.verts["sec"], lv.verts["tex"], lv.idx,
                    lv.tverts["pos"], lv.tverts["sec"], lv.tverts["tex"], lv.tidx),
        "normals": _h(lv.verts["nrm"], lv.tverts["nrm"]),
    }
