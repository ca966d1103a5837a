"""Sharded runs (include/vxb200.h: vxb_polygonize_sharded, vxb_cube_*) on ONE GPU.

`world` contexts of one process play the ranks: each polygonizes its z-slab, the one exchange step (material pages of
the last nested level) is a device copy, and the merged result must be bit-identical to the unsharded run - and thus
to the reference.  The cube test backs every slab with its own cuMemCreate allocation and maps the "peers'" slabs
through exported file descriptors, exactly what the ranks of a torchrun job do with each other's HBM."""
import os

import numpy as np
import pytest

import compare
import grids

pytestmark = pytest.mark.gpu


def _views(x, world):
    import torch
    from voxels_b200 import capi
    dev = torch.device("cuda", 0)
    return (torch.as_tensor(capi.DevicePointer(x.pages, x.pages_bytes), device=dev),
            torch.as_tensor(capi.DevicePointer(x.valid, x.valid_bytes), device=dev))


def run_virtual_ranks(contexts, world, flags=0, planes=None):
    """contexts[r] already sees the whole cube.  Returns the merged result."""
    import torch
    from voxels_b200 import capi
    from voxels_b200.dist import uniform_planes
    n = contexts[0].n
    pb = uniform_planes(n, world) if planes is None else planes
    for attempt in range(4):
        for r, c in enumerate(contexts):
            c.polygonize_sharded(r, world, 0, flags, planes)
        torch.cuda.synchronize()
        xs = [c.shard_exchange_info(r, world, planes) for r, c in enumerate(contexts)]
        if world > 1:
            views = [_views(x, world) for x in xs]
            x = xs[0]
            for kind, unit in ((0, 8192), (1, 1)):
                for src in range(world):
                    lo = pb[src] // x.layer_planes * x.layer_blocks * unit
                    hi = pb[src + 1] // x.layer_planes * x.layer_blocks * unit
                    for dst in range(world):
                        if src != dst:
                            views[dst][kind][lo:hi].copy_(views[src][kind][lo:hi])
            torch.cuda.synchronize()
        rcs = [c.polygonize_sharded(r, world, 1, flags, planes) for r, c in enumerate(contexts)]
        if not any(rcs):
            return capi.merge_results([c.download() for c in contexts])
    raise AssertionError("arenas kept overflowing")


def assert_same(a, b, levels):
    problems = []
    for l in range(levels):
        problems += compare.level_diff(a.level(l), b.level(l), "L%d" % l)
    if not np.array_equal(a.stats, b.stats):
        problems.append("stats differ: %s vs %s" % (a.stats, b.stats))
    assert not problems, "\n".join(problems[:10])


@pytest.mark.parametrize("name,world,planes", [("hostile128", 2, None), ("hostile128", 4, None), ("sphere128", 2, None), ("noise64", 2, None),
                                               ("hostile64", 2, None), ("hostile128", 3, [0, 32, 96, 128]), ("sphere128", 2, [0, 96, 128])])
def test_virtual_ranks_equal_single_run(gpu_context, name, world, planes):
    import voxels_b200
    dist, mat, blend = (grids.MEDIUM.get(name) or grids.SMALL[name])()
    n = dist.shape[0]
    gpu_context.set_materials(None, None)
    gpu_context.upload_dense(dist, mat, blend)
    info = gpu_context.polygonize()
    single = gpu_context.download()
    d, m, b = gpu_context.device_pointers()
    contexts = [voxels_b200.Context(0) for _ in range(world)]
    try:
        for c in contexts:
            c.set_device_grid(n, d, m, b)
        merged = run_virtual_ranks(contexts, world, planes=planes)
        assert_same(single, merged, info.levels_total)
        # every block exactly once, ids are the full-run ids
        assert np.array_equal(single.records["id"], merged.records["id"])
    finally:
        for c in contexts:
            c.close()


def test_sharded_against_reference(reference, gpu_context):
    import voxels_b200
    dist, mat, blend = grids.MEDIUM["hostile128"]()
    g = reference.grid_from_dense(dist, mat, blend)
    s, _ = reference.polygonize(g)
    gpu_context.upload_dense(dist, mat, blend)
    d, m, b = gpu_context.device_pointers()
    contexts = [voxels_b200.Context(0) for _ in range(4)]
    try:
        for c in contexts:
            c.set_device_grid(128, d, m, b)
        merged = run_virtual_ranks(contexts, 4)
        problems = []
        for l in range(reference.surface_levels(s)):
            problems += compare.level_diff(reference.surface_level(s, l), merged.level(l), "L%d" % l)
        if not np.array_equal(reference.surface_stats(s), merged.stats):
            problems.append("stats differ")
        assert not problems, "\n".join(problems[:10])
    finally:
        for c in contexts:
            c.close()
        reference.surface_destroy(s)
        reference.grid_destroy(g)


@pytest.mark.parametrize("world,planes", [(1, None), (2, None), (4, None), (3, [0, 64, 192, 256])])
def test_cube_of_mapped_slabs(gpu_context, world, planes):
    """The VMM cube: one virtual range per volume, every slab its own physical allocation, peers imported by descriptor."""
    import torch
    import voxels_b200
    from voxels_b200 import capi, synth
    n = 256
    dev = torch.device("cuda", 0)
    dist, mat, blend = synth.terrain(n, dev)
    gpu_context.set_materials(None, None)
    gpu_context.set_device_grid(n, dist.data_ptr(), mat.data_ptr(), blend.data_ptr(), keep=(dist, mat, blend))
    info = gpu_context.polygonize()
    single = gpu_context.download()
    contexts = [voxels_b200.Context(0) for _ in range(world)]
    try:
        from voxels_b200.dist import uniform_planes
        pb = uniform_planes(n, world) if planes is None else planes
        for r, c in enumerate(contexts):
            c.cube_create(n, r, world, planes)
        for r, c in enumerate(contexts):
            fds = [c.cube_export(ch) for ch in range(3)]
            for p, peer in enumerate(contexts):
                if p != r:
                    for ch, fd in enumerate(fds):
                        peer.cube_import(r, ch, fd)
            for fd in fds:
                os.close(fd)
        for r, c in enumerate(contexts):
            pd, pm, pbl, size = c.cube_slab()
            assert size == n * n * (pb[r + 1] - pb[r])
            for ptr, src in ((pd, dist), (pm, mat), (pbl, blend)):
                torch.as_tensor(capi.DevicePointer(ptr, size), device=dev).copy_(src[pb[r]:pb[r + 1]].reshape(-1).view(torch.uint8))
        torch.cuda.synchronize()
        merged = run_virtual_ranks(contexts, world, planes=planes)
        assert_same(single, merged, info.levels_total)
    finally:
        for c in contexts:
            c.close()
