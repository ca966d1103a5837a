"""Synthetic grids for the parity tests (numpy; small sizes).  All arrays are indexed [z, y, x]."""
import numpy as np


def _hash3(x, y, z, seed):
    h = (x.astype(np.uint64) * np.uint64(73856093)) ^ (y.astype(np.uint64) * np.uint64(19349663)) ^ \
        (z.astype(np.uint64) * np.uint64(83492791)) ^ np.uint64(seed * 2654435761 & 0xFFFFFFFF)
    h &= np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(13); h = (h * np.uint64(0x5BD1E995)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(15); h = (h * np.uint64(0x27D4EB2F)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(16)
    return h


def sphere_floats(n, center=None, radius=None):
    c = (n / 2.0,) * 3 if center is None else center
    r = 0.3 * n if radius is None else radius
    z, y, x = np.meshgrid(np.arange(n, dtype=np.float32), np.arange(n, dtype=np.float32), np.arange(n, dtype=np.float32), indexing="ij")
    d = np.sqrt((x - c[0]) ** 2 + (y - c[1]) ** 2 + (z - c[2]) ** 2) - np.float32(r)
    return np.clip(d, -100, 100).astype(np.float32)


def hostile(n, seed=7, vmax=3, materials=3):
    """wavy terrain + hash noise near the surface + forced zeros, hash materials/blends.
    Returns already-quantised int8 distances."""
    z, y, x = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    d = z - (n / 2.0 + 6 * np.sin(0.21 * x) * np.cos(0.17 * y) + 3 * np.sin(0.05 * (x + y)))
    q = (np.sign(d) * np.ceil(np.abs(d))).astype(np.int64)
    h = _hash3(x, y, z, seed)
    near = np.abs(d) < 6
    q = np.where(near, q + (h % np.uint64(2 * vmax + 1)).astype(np.int64) - vmax, q)
    q = np.where(near & ((h >> np.uint64(8)) % np.uint64(4) == 0), 0, q)
    lim = max(4, vmax + 1)
    q = np.clip(q, -lim, lim).astype(np.int8)
    hm = _hash3(x // 5, y // 5, z // 3, seed + 1)
    mat = (hm % np.uint64(materials)).astype(np.uint8)
    blend = ((h >> np.uint64(16)) & np.uint64(0xFF)).astype(np.uint8)
    return np.ascontiguousarray(q), np.ascontiguousarray(mat), np.ascontiguousarray(blend)


def noise_full(n, seed=3, lo=-5, hi=5):
    """Every voxel random in [lo, hi]: (almost) every cell non-trivial, worst case for counts."""
    z, y, x = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    h = _hash3(x, y, z, seed)
    q = (h % np.uint64(hi - lo + 1)).astype(np.int64) + lo
    mat = ((h >> np.uint64(10)) % np.uint64(4)).astype(np.uint8)
    blend = ((h >> np.uint64(20)) & np.uint64(0xFF)).astype(np.uint8)
    return np.ascontiguousarray(q.astype(np.int8)), np.ascontiguousarray(mat), np.ascontiguousarray(blend)
