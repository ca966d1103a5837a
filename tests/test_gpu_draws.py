"""Consumer-side packaging (include/vxb200.h "consumer side"): the LOD cut and the indirect draw
lists built on the GPU, checked against a direct numpy evaluation of the renderer contract in doc_source/Rendering.md."""
import numpy as np
import pytest

from voxels_b200 import capi

pytestmark = pytest.mark.gpu


def _stops(level, b, cam, base, top):
    if level == top:
        return True
    m = 16 << level
    c = (np.array([b[0], b[2], b[1]], np.float64) + 0.5) * m          # output axes (x, z, y)
    return float(((c - cam) ** 2).sum()) >= float((base * (1 << level)) ** 2)


def _selected(level, b, cam, base, top):
    if level > 0 and not _stops(level, b, cam, base, top):
        return False
    return not any(_stops(l, tuple(v >> (l - level) for v in b), cam, base, top) for l in range(level + 1, top + 1))


def _finer(level, b, cam, base, top):
    return level > 0 and not any(_stops(l, tuple(v >> (l - level) for v in b), cam, base, top) for l in range(level, top + 1))


@pytest.mark.parametrize("camera,base", [((40.0, 70.0, 50.0), 48.0), ((128.0, 64.0, 128.0), 20.0), ((-50.0, 300.0, 10.0), 64.0)])
def test_lod_cut_and_draw_lists(gpu_context, camera, base):
    n = 128
    ctx = gpu_context
    ctx.set_materials(None, None)
    ctx.fill(n, capi.Surface.terrain(n))
    info = ctx.polygonize()
    res = ctx.download()
    recs = res.records
    top = info.levels_total - 1
    cam = np.array(camera, np.float64)
    rc, ri, tc, ti = ctx.select_lod(camera, base)
    # expected selection, block by block
    want = {}
    for r in recs:
        level = int(r["level"]); nb = (n // 16) >> level; c = int(r["coord_id"])
        b = (c % nb, (c // nb) % nb, c // (nb * nb))
        if not _selected(level, b, cam, base, top):
            continue
        adj = 0
        if level > 0:
            for f, d in enumerate([(0, 0, -1), (0, -1, 0), (-1, 0, 0), (0, 0, 1), (0, 1, 0), (1, 0, 0)]):
                nbr = (b[0] + d[0], b[1] + d[1], b[2] + d[2])
                if min(nbr) < 0 or max(nbr) >= nb:
                    continue
                if _finer(level, nbr, cam, base, top):
                    adj |= 1 << f
        want[int(r["id"])] = (r, adj)
    assert len(rc) == len(want) > 0
    got_ids = set(int(i) for i in ri["block_id"])
    assert got_ids == set(want)
    assert np.array_equal(rc["first_instance"], np.arange(len(rc)))        # first_instance indexes the info array
    trans_expected = 0
    for cmd, inf in zip(rc, ri):
        r, adj = want[int(inf["block_id"])]
        assert int(inf["block_adj"]) == adj and int(inf["level"]) == int(r["level"]) and int(inf["face"]) == 0xFFFFFFFF
        assert (int(cmd["index_count"]), int(cmd["first_index"]), int(cmd["base_vertex"]), int(cmd["instance_count"])) == \
               (int(r["index_count"]), int(r["index_offset"]), int(r["vertex_offset"]), 1)
        trans_expected += sum(1 for f in range(6) if (adj >> f) & 1 and r["trans_index_count"][f])
    assert len(tc) == trans_expected
    for cmd, inf in zip(tc, ti):
        r, adj = want[int(inf["block_id"])]
        f = int(inf["face"])
        assert (adj >> f) & 1 and int(inf["block_adj"]) == adj
        assert (int(cmd["index_count"]), int(cmd["first_index"]), int(cmd["base_vertex"])) == \
               (int(r["trans_index_count"][f]), int(r["trans_index_offset"][f]), int(r["trans_vertex_offset"][f]))
    # the cut is a cut: no drawn block has a drawn ancestor
    drawn = {(int(r["level"]), int(r["coord_id"])) for r, _ in want.values()}
    for level, c in drawn:
        nb = (n // 16) >> level
        b = (c % nb, (c // nb) % nb, c // (nb * nb))
        for l in range(level + 1, top + 1):
            a = tuple(v >> (l - level) for v in b); nba = (n // 16) >> l
            assert (l, (a[2] * nba + a[1]) * nba + a[0]) not in drawn
    if len(tc):
        assert ctx.device_arenas()[2] != 0
