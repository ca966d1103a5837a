"""Host side of the drop-in's full Execute that needs no GPU: the gather of the grid store's compressed blocks into one blob
(polygonizer_host.cpp: blobLayout / packBlocks, exported as voxels_b200_pack_grid) must produce exactly the bytes
Grid::PackForSave returns (reference src/VoxelGrid.cpp:269-315) - the form vxb_grid_upload_packed[_streamed] decodes."""
import ctypes
import os

import numpy as np
import pytest

import grids
import harness


@pytest.fixture(scope="module")
def dropin():
    if not os.path.exists(harness.B200_LIB):
        pytest.skip("drop-in harness not built")
    return harness.load(harness.B200_LIB)


@pytest.fixture(scope="module")
def pack_grid():
    lib = ctypes.CDLL(os.path.join(harness.REPO, "voxels_b200", "lib", "libvoxels_b200.so"))
    lib.voxels_b200_pack_grid.restype = ctypes.c_size_t
    lib.voxels_b200_pack_grid.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
    return lib.voxels_b200_pack_grid


@pytest.mark.parametrize("name", ["hostile64", "noise32", "sphere64", "positive_noise32", "plane32"])
def test_gather_equals_pack_for_save(dropin, pack_grid, name):
    import voxels_b200
    dist, mat, blend = grids.SMALL[name]()
    g = dropin.grid_from_dense(dist, mat, blend)
    try:
        want = dropin.grid_pack(g)
        size = pack_grid(g, None, 0)
        assert size == len(want)
        out = np.zeros(size + 8, np.uint8)
        out[size:] = 0xAB
        assert pack_grid(g, out.ctypes.data, size) == size
        assert np.array_equal(out[:size], want)
        assert (out[size:] == 0xAB).all()          # nothing past the end
        assert pack_grid(g, out.ctypes.data, size - 1) == size   # too small: size reported, nothing written
        # and the library's own packer of dense volumes makes the same bytes
        assert np.array_equal(voxels_b200.pack_dense(dist, mat, blend), want)
    finally:
        dropin.grid_destroy(g)


def test_gather_thread_override(dropin, pack_grid, monkeypatch):
    """VXB200_PACK_THREADS is read once per process; whatever the team size, the bytes are the same (static schedule over blocks)."""
    dist, mat, blend = grids.SMALL["hostile64"]()
    g = dropin.grid_from_dense(dist, mat, blend)
    try:
        want = dropin.grid_pack(g)
        for _ in range(3):
            out = np.empty(len(want), np.uint8)
            assert pack_grid(g, out.ctypes.data, len(out)) == len(want)
            assert np.array_equal(out, want)
    finally:
        dropin.grid_destroy(g)
