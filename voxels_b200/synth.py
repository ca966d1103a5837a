"""Synthetic workloads of BASELINE.json / SURVEY.md section 8(d), generated with torch (device plumbing only:
this produces INPUT DATA outside every timed region; nothing here is part of the polygonizer).

terrain(n): seeded Perlin terrain, 4 materials, smooth blends, quantised exactly like the reference's grid
constructor (VoxelGrid.cpp:37-50: round away from zero, clamp to +-4).  The same bytes feed the CUDA path
and the CPU reference, so the generator's own float rounding never enters a parity comparison.
"""
import numpy as np
import torch

_GRAD3 = [[1, 1, 0], [-1, 1, 0], [1, -1, 0], [-1, -1, 0], [1, 0, 1], [-1, 0, 1], [1, 0, -1], [-1, 0, -1],
          [0, 1, 1], [0, -1, 1], [0, 1, -1], [0, -1, -1], [1, 1, 0], [-1, 1, 0], [0, -1, 1], [0, -1, -1]]


def _fade(t):
    return t * t * t * (t * (t * 6 - 15) + 10)


class Perlin:
    """Classic gradient noise; permutation table from numpy RandomState(seed) so it is identical everywhere."""

    def __init__(self, seed, device):
        perm = np.random.RandomState(seed).permutation(256)
        self.perm = torch.tensor(np.concatenate([perm, perm, perm]), device=device, dtype=torch.long)
        self.grad = torch.tensor(_GRAD3, device=device, dtype=torch.float32)

    def noise3(self, x, y, z):
        """x: [1,1,X], y: [1,Y,1], z: [Z,1,1] float tensors -> [Z,Y,X]."""
        xf, yf, zf = torch.floor(x), torch.floor(y), torch.floor(z)
        xi, yi, zi = xf.long() & 255, yf.long() & 255, zf.long() & 255
        dx, dy, dz = x - xf, y - yf, z - zf
        u, v, w = _fade(dx), _fade(dy), _fade(dz)
        p = self.perm
        out = None
        for cz in (0, 1):
            for cy in (0, 1):
                for cx in (0, 1):
                    h = p[p[p[xi + cx] + yi + cy] + zi + cz] & 15
                    g = self.grad[h]
                    dot = g[..., 0] * (dx - cx) + g[..., 1] * (dy - cy) + g[..., 2] * (dz - cz)
                    wgt = (u if cx else 1 - u) * (v if cy else 1 - v) * (w if cz else 1 - w)
                    out = dot * wgt if out is None else out + dot * wgt
        return out

    def noise2(self, x, y):
        """x: [1,X], y: [Y,1] -> [Y,X] (the z = 0.5 slice of the 3-D noise)."""
        z = torch.full((1, 1, 1), 0.5, device=x.device)
        return self.noise3(x.reshape(1, 1, -1), y.reshape(1, -1, 1), z)[0]

    def fbm2(self, x, y, octaves=5, lacunarity=2.0, gain=0.5):
        total, amp, freq = 0.0, 1.0, 1.0
        for _ in range(octaves):
            total = total + amp * self.noise2(x * freq, y * freq)
            amp *= gain
            freq *= lacunarity
        return total


def quantise(d):
    """float SDF -> int8 the way the reference does: sign(v)*ceil(|v|), then clamp to [-4, 4]."""
    return torch.clamp(torch.sign(d) * torch.ceil(torch.abs(d)), -4, 4).to(torch.int8)


def terrain(n, device="cpu", seed=1234, origin=(0, 0), z_chunk=16, z_range=None, out=None):
    """(dist int8, mat uint8, blend uint8), each [n, n, n] indexed [z, y, x], on `device`.
    origin shifts the (x, y) window so different ranks get different tiles of one endless terrain.
    z_range=(z0, z1) produces only those planes of the same n^3 terrain ([z1-z0, n, n]: the slab of one rank of a
    sharded run); out=(dist, mat, blend) writes into existing tensors of that shape."""
    dev = torch.device(device)
    pn = Perlin(seed, dev)
    xs = torch.arange(n, device=dev, dtype=torch.float32) + float(origin[0])
    ys = torch.arange(n, device=dev, dtype=torch.float32) + float(origin[1])
    X2, Y2 = xs.reshape(1, -1), ys.reshape(-1, 1)
    height = 0.5 * n + 0.18 * n * pn.fbm2(X2 / n * 4.0, Y2 / n * 4.0)                   # [Y, X]
    h1 = 0.45 * n + 0.03 * n * pn.noise2(X2 / 37.0, Y2 / 37.0)
    h2 = 0.60 * n + 0.03 * n * pn.noise2(X2 / 53.0 + 7.7, Y2 / 53.0 + 3.3)
    zlo, zhi = (0, n) if z_range is None else z_range
    if out is None:
        dist = torch.empty((zhi - zlo, n, n), dtype=torch.int8, device=dev)
        mat = torch.empty((zhi - zlo, n, n), dtype=torch.uint8, device=dev)
        blend = torch.empty((zhi - zlo, n, n), dtype=torch.uint8, device=dev)
    else:
        dist, mat, blend = out
    X3, Y3 = xs.reshape(1, 1, -1), ys.reshape(1, -1, 1)
    for z0 in range(zlo, zhi, z_chunk):
        z1 = min(zhi, z0 + z_chunk)
        zs = torch.arange(z0, z1, device=dev, dtype=torch.float32).reshape(-1, 1, 1)
        d = zs - height.unsqueeze(0) + 6.0 * pn.noise3(X3 / 24.0, Y3 / 24.0, zs / 24.0)
        dist[z0 - zlo:z1 - zlo] = quantise(torch.clamp(d, -100, 100))
        m = torch.where(zs < h1.unsqueeze(0), 0, torch.where(zs < h2.unsqueeze(0), 1, 2))
        ore = pn.noise3(X3 / 48.0 + 11.1, Y3 / 48.0 + 5.5, zs / 48.0 + 2.2) > 0.35
        mat[z0 - zlo:z1 - zlo] = torch.where(ore, 3, m).to(torch.uint8)
        t = torch.clamp((zs - h1.unsqueeze(0)) / (h2 - h1).unsqueeze(0).clamp(min=1.0), 0, 1)
        blend[z0 - zlo:z1 - zlo] = (255.0 * (t * t * (3 - 2 * t))).to(torch.uint8)
    return dist, mat, blend


def sphere(n, device="cpu", center=None, radius=None):
    """Config 1: d = |p - c| - r (c = n/2, r = 0.3 n), one material."""
    dev = torch.device(device)
    c = (n / 2.0,) * 3 if center is None else center
    r = 0.3 * n if radius is None else radius
    a = torch.arange(n, device=dev, dtype=torch.float32)
    d = torch.sqrt((a.reshape(1, 1, -1) - c[0]) ** 2 + (a.reshape(1, -1, 1) - c[1]) ** 2 + (a.reshape(-1, 1, 1) - c[2]) ** 2) - r
    z = torch.zeros((n, n, n), dtype=torch.uint8, device=dev)
    return quantise(torch.clamp(d, -100, 100)), z, z.clone()
