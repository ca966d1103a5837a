"""Named seeded test grids shared by the CPU and GPU parity tests: name -> (dist, mat, blend) int8/uint8 [z,y,x]."""
import numpy as np

import gridgen


def _quantise(d):
    q = np.sign(d) * np.ceil(np.abs(d))  # reference `round` (VoxelGrid.cpp:37-40) + clamp +-4 (:42-50)
    return np.clip(q, -4, 4).astype(np.int8)


def sphere(n):
    d = _quantise(gridgen.sphere_floats(n))
    z = np.zeros((n, n, n), np.uint8)
    return d, z, z.copy()


def two_material_plane(n):
    z, y, x = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    d = _quantise(0.37 * (x - n / 2.0) + 0.61 * (y - n / 2.0) + 0.7 * (z - n / 2.0) + 0.25)
    mat = np.where(x + y > n, 1, 2).astype(np.uint8)
    blend = ((x * 3 + y * 5 + z * 7) % 256).astype(np.uint8)
    return np.ascontiguousarray(d), np.ascontiguousarray(mat), np.ascontiguousarray(blend)


SMALL = {
    "sphere64": lambda: sphere(64),
    "plane32": lambda: two_material_plane(32),
    "hostile64": lambda: gridgen.hostile(64),
    "hostile64_big_values": lambda: gridgen.hostile(64, seed=5, vmax=8, materials=5),
    "noise32": lambda: gridgen.noise_full(32),
    "noise16": lambda: gridgen.noise_full(16, seed=2),
    "positive_noise32": lambda: gridgen.noise_full(32, seed=4, lo=1, hi=4),
}
MEDIUM = {
    "hostile128": lambda: gridgen.hostile(128, seed=11),
    "sphere128": lambda: sphere(128),
    "noise64": lambda: gridgen.noise_full(64, seed=9, lo=-3, hi=2),
}
