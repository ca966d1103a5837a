"""CPU-side checks of the product's C ABI: the library loads, exports every symbol include/vxb200.h declares,
the Python struct mirrors match the header, and - without a CUDA device - every entry point fails loudly
(there is no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

import voxels_b200
from voxels_b200 import capi

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(REPO, "include", "vxb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vxb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(voxels_b200.library_path())
    names = declared_symbols()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), "include/vxb200.h declares %s but libvxb200.so does not export it" % name
    assert sorted(capi.EXPORTED_SYMBOLS) == names, "voxels_b200/capi.py binds a different set than the header declares"


def test_struct_mirrors_match_header_sizes():
    assert capi.RECORD_DTYPE.itemsize == 128 and capi.VERTEX_DTYPE.itemsize == 48
    assert C.sizeof(capi.ResultInfo) == 4 * 4 + 8 * 8 + 20 * 4 + 8 * 4 + 4 + 4


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(voxels_b200.VxbError) as e:
        voxels_b200.Context(0)
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)


def test_dropin_exports_the_reference_symbols():
    """libvoxels_b200.so must define what the reference's TransVoxelImpl.cpp defines for the public headers."""
    import subprocess
    path = os.path.join(REPO, "voxels_b200", "lib", "libvoxels_b200.so")
    if not os.path.exists(path):
        pytest.skip("drop-in not built (needs the reference headers)")
    out = subprocess.run(["nm", "-D", "--defined-only", "-C", path], capture_output=True, text=True, check=True).stdout
    for sym in ["Voxels::Polygonizer::Polygonizer()", "Voxels::Polygonizer::~Polygonizer()",
                "Voxels::Polygonizer::Execute(Voxels::Grid const&, Voxels::MaterialMap const*, Voxels::Modification*)",
                "Voxels::Modification::Create()", "Voxels::Modification::~Modification()", "Voxels::PolygonSurface::INVALID_ID",
                "InitializeVoxels", "DeinitializeVoxels", "GetBuildVersion", "Voxels::Grid::Create("]:
        assert sym in out, "missing symbol: " + sym


def test_tables_checksum():
    """The generated Transvoxel tables (build/gen, from the reference's Transvoxel.inl) keep their checksum."""
    path = os.path.join(REPO, "build", "gen", "vxb_tables_data.h")
    if not os.path.exists(path):
        pytest.skip("tables not generated")
    text = open(path).read()
    assert "VXB_TABLES_FNV1A 0x83E0932026BB2FEEull" in text
    assert "Eric Lengyel's Transvoxel Algorithm" in text and "http://www.terathon.com/voxels/" in text


def test_pack_dense_matches_reference_bytes(reference):
    """vxb_pack_dense (host helper of the C ABI) writes exactly the bytes Grid::PackForSave produces."""
    import numpy as np
    import grids
    for name in ("hostile64", "positive_noise32", "plane32"):
        dist, mat, blend = grids.SMALL[name]()
        g = reference.grid_from_dense(dist, mat, blend)
        blob = reference.grid_pack(g)
        reference.grid_destroy(g)
        assert np.array_equal(blob, voxels_b200.pack_dense(dist, mat, blend)), name
