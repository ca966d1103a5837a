"""The built-in procedural surfaces (voxels_b200/csrc/vxb_surfaces.h) on the host: the dense generator used for full-size
inputs must produce exactly the bytes the UNMODIFIED reference grid constructor makes of the same surface
(Grid::Create(..., &surface): VoxelGrid.cpp:79-132 - sample, round away from zero, clamp to +-4)."""
import numpy as np
import pytest

from voxels_b200 import capi


@pytest.mark.parametrize("make,n", [(lambda n: capi.Surface.sphere((n / 2, n / 2, n / 2), 0.3 * n, 2, 17), 64),
                                    (lambda n: capi.Surface.plane((0.37, 0.61, 0.7), 40.25, 1, 200), 32),
                                    (lambda n: capi.Surface.terrain(n), 64),
                                    (lambda n: capi.Surface.terrain(n, origin=(64, 0), seed=7), 32)])
def test_dense_generator_equals_reference_constructor(reference, make, n):
    s = make(n)
    g = reference.grid_create_builtin(n, s)
    want = reference.grid_to_dense(g)
    reference.grid_destroy(g)
    got = reference.builtin_dense(n, s)
    for a, b, what in zip(want, got, ("distance", "material", "blend")):
        assert np.array_equal(a, b), what
    if s.kind == 2:
        d, m, b = got
        assert d.min() == -4 and d.max() == 4 and len(np.unique(m)) >= 3 and b.max() > b.min()   # a real terrain, not a constant
        frac = np.mean(np.abs(d.astype(np.int32)) < 4)
        assert 0.01 < frac < 0.5


def test_terrain_tiles_continue_each_other(reference):
    """origin shifts the (x, y) window of one endless terrain: the tile at origin (32, 0) repeats the right half of (0, 0)'s."""
    a = reference.builtin_dense(64, capi.Surface.terrain(256))
    b = reference.builtin_dense(64, capi.Surface.terrain(256, origin=(32, 0)))
    for x, y in zip(a, b):
        assert np.array_equal(x[:, :, 32:], y[:, :, :32])
