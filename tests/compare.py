"""Comparison rules of SURVEY.md section 8(c): which fields must be bit-exact, which get a tolerance."""
import numpy as np

NORMAL_TOL = 1e-5


def _verts_diff(a, b, what):
    msgs = []
    if len(a) != len(b):
        return ["%s: vertex count %d != %d" % (what, len(a), len(b))]
    if len(a) == 0:
        return msgs
    for field in ("pos", "sec"):
        # positions are exact dyadic rationals -> compare bit patterns (sec.w carries flag bits)
        ua, ub = a[field].view(np.uint32), b[field].view(np.uint32)
        if not np.array_equal(ua, ub):
            bad = np.nonzero((ua != ub).any(axis=1))[0]
            msgs.append("%s: %s differs at %d vertices, first %d: %s vs %s" % (what, field, len(bad), bad[0], a[field][bad[0]], b[field][bad[0]]))
    if not np.array_equal(a["tex"], b["tex"]):
        bad = np.nonzero((a["tex"] != b["tex"]).any(axis=1))[0]
        msgs.append("%s: texture bytes differ at %d vertices, first %d: %s vs %s" % (what, len(bad), bad[0], a["tex"][bad[0]], b["tex"][bad[0]]))
    dn = np.abs(a["nrm"] - b["nrm"])
    if dn.size and dn.max() > NORMAL_TOL:
        bad = np.nonzero((dn > NORMAL_TOL).any(axis=1))[0]
        msgs.append("%s: normals differ (max %g) at %d vertices, first %d: %s vs %s" % (what, dn.max(), len(bad), bad[0], a["nrm"][bad[0]], b["nrm"][bad[0]]))
    return msgs


def level_diff(a, b, what="level"):
    """List of human-readable mismatches between two LevelDumps (empty = parity)."""
    msgs = []
    if len(a.rows) != len(b.rows):
        msgs.append("%s: block count %d != %d" % (what, len(a.rows), len(b.rows)))
        return msgs
    for f in ("id", "min", "max", "nv", "ni", "tnv", "tni"):
        if not np.array_equal(a.rows[f], b.rows[f]):
            ne = a.rows[f] != b.rows[f]
            bad = np.nonzero(ne if ne.ndim == 1 else ne.any(axis=1))[0]
            msgs.append("%s: block table field %s differs in %d blocks, first %d: %s vs %s" % (what, f, len(bad), bad[0], a.rows[f][bad[0]], b.rows[f][bad[0]]))
    if msgs:
        return msgs
    msgs += _verts_diff(a.verts, b.verts, what + " regular")
    if not np.array_equal(a.idx, b.idx):
        bad = np.nonzero(a.idx != b.idx)[0]
        msgs.append("%s: indices differ at %d positions, first %d" % (what, len(bad), bad[0]))
    msgs += _verts_diff(a.tverts, b.tverts, what + " transition")
    if not np.array_equal(a.tidx, b.tidx):
        bad = np.nonzero(a.tidx != b.tidx)[0]
        msgs.append("%s: transition indices differ at %d positions, first %d" % (what, len(bad), bad[0]))
    return msgs


def normals_max_ulp(a, b):
    worst = 0
    for va, vb in ((a.verts, b.verts), (a.tverts, b.tverts)):
        if len(va) == 0 or len(va) != len(vb):
            continue
        ia = va["nrm"].view(np.int32).astype(np.int64); ib = vb["nrm"].view(np.int32).astype(np.int64)
        worst = max(worst, int(np.abs(ia - ib).max()))
    return worst
