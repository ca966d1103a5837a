"""GPU parity: the CUDA path (through the C ABI, include/vxb200.h) against the unmodified reference.

Bit-exact: block ids/order/corners, counts, indices, transition indices, texture bytes, flag bits,
positions, statistics.  Normals: <= 1e-5 (contract of BASELINE.json's north_star), 0 ULP expected."""
import numpy as np
import pytest

import compare
import grids

pytestmark = pytest.mark.gpu


def run_both(reference, ctx, dist, mat, blend, table=None, valid=None, max_levels=0, flags=0):
    g = reference.grid_from_dense(dist, mat, blend)
    s, _ = reference.polygonize(g, material_table=table, valid_mask=valid)
    ctx.set_materials(table, valid)
    ctx.upload_dense(dist, mat, blend)
    info = ctx.polygonize(max_levels, flags)
    res = ctx.download()
    problems = []
    levels = reference.surface_levels(s)
    assert info.levels_total == levels
    ncmp = levels if max_levels == 0 else min(levels, max_levels)
    for l in range(ncmp):
        problems += compare.level_diff(reference.surface_level(s, l), res.level(l), "L%d" % l)
    if max_levels == 0:
        rs = reference.surface_stats(s)
        if not np.array_equal(rs, res.stats):
            problems.append("stats differ: ref %s gpu %s" % (rs, res.stats))
    reference.surface_destroy(s)
    reference.grid_destroy(g)
    return problems, res


@pytest.mark.parametrize("name", sorted(grids.SMALL))
def test_small_grids_all_levels(reference, gpu_context, name):
    dist, mat, blend = grids.SMALL[name]()
    problems, res = run_both(reference, gpu_context, dist, mat, blend)
    assert not problems, "\n".join(problems[:10])


@pytest.mark.parametrize("name", sorted(grids.MEDIUM))
def test_medium_grids_all_levels(reference, gpu_context, name):
    dist, mat, blend = grids.MEDIUM[name]()
    problems, res = run_both(reference, gpu_context, dist, mat, blend)
    assert not problems, "\n".join(problems[:10])


def test_material_table_and_invalid_ids(reference, gpu_context):
    dist, mat, blend = grids.SMALL["hostile64"]()
    table = (np.arange(256 * 6) * 7 % 251).astype(np.uint8)
    valid = np.ones(256, np.uint8); valid[1] = 0  # GetMaterial(1) -> nullptr: textures stay zero (TransVoxelImpl.cpp:1364-1368)
    problems, res = run_both(reference, gpu_context, dist, mat, blend, table=table, valid=valid)
    assert not problems, "\n".join(problems[:10])


def test_level0_only_regular_cells(reference, gpu_context):
    """BASELINE config 2 shape: single LOD, regular cells only."""
    import voxels_b200
    dist, mat, blend = grids.MEDIUM["hostile128"]()
    problems, res = run_both(reference, gpu_context, dist, mat, blend, max_levels=1, flags=voxels_b200.FLAG_NO_TRANSITIONS)
    assert not problems, "\n".join(problems[:10])


def test_block_upload_equals_dense_upload(gpu_context):
    dist, mat, blend = grids.SMALL["hostile64"]()
    n = dist.shape[0]; nb = n // 16

    def to_blocks(a):
        return np.ascontiguousarray(a.reshape(nb, 16, nb, 16, nb, 16).transpose(0, 2, 4, 1, 3, 5)).reshape(-1)

    gpu_context.set_materials(None, None)
    gpu_context.upload_dense(dist, mat, blend)
    gpu_context.polygonize()
    a = gpu_context.download()
    gpu_context.upload_blocks(n, to_blocks(dist), to_blocks(mat), to_blocks(blend))
    gpu_context.polygonize()
    b = gpu_context.download()
    for l in range(a.info.levels_total):
        assert not compare.level_diff(a.level(l), b.level(l), "L%d" % l)


def test_arena_growth_retry(reference, gpu_context):
    """Tiny initial arenas: the run must detect the overflow, grow and still match."""
    import voxels_b200
    ctx = voxels_b200.Context(0)
    try:
        ctx.set_capacity(1024, 1024, 64, 64)
        dist, mat, blend = grids.SMALL["hostile64"]()
        problems, res = run_both(reference, ctx, dist, mat, blend)
        assert not problems, "\n".join(problems[:10])
    finally:
        ctx.close()


def test_repeatable(gpu_context):
    dist, mat, blend = grids.SMALL["hostile64"]()
    gpu_context.set_materials(None, None)
    gpu_context.upload_dense(dist, mat, blend)
    gpu_context.polygonize(); a = gpu_context.download()
    gpu_context.polygonize(); b = gpu_context.download()
    for l in range(a.info.levels_total):
        assert not compare.level_diff(a.level(l), b.level(l), "L%d" % l)
    assert np.array_equal(a.stats, b.stats)
