// Built-in procedural surfaces: the device-side analogue of a client's Voxels::VoxelSurface (reference
// include/VoxelSurface.h:35-40).  Written once and compiled
//   * by nvcc as __device__ code for the fill / edit kernels (vxb_grid.cuh: Grid::Create src/VoxelGrid.cpp:79-132,
//     VoxelGrid::InjectSurface :388-488), and
//   * by g++ as plain inline C++ for the VoxelSurface adapter of the test harness (tests/harness/vxh_capi.cpp), which
//     feeds the UNMODIFIED reference grid store, so that both sides evaluate bit-identical floats.
// Floating point: +, -, *, /, sqrt, floor only, in a fixed order, compiled without FMA contraction (-fmad=false for
// nvcc; g++ -O2 -msse2 has no FMA), IEEE division and square root on both sides.
#pragma once
#include <stdint.h>

#include "../../include/vxb200.h" // vxb_surface

#if defined(__CUDACC__)
#define VXS_FN __device__ __forceinline__
#define VXS_SQRT(x) __fsqrt_rn(x)
#define VXS_DIV(a, b) __fdiv_rn((a), (b))
#define VXS_FLOOR(x) floorf(x)
#else
#include <cmath>
#define VXS_FN inline
#define VXS_SQRT(x) std::sqrt(x)
#define VXS_DIV(a, b) ((a) / (b))
#define VXS_FLOOR(x) std::floor(x)
#endif

// permutation table of the gradient noise: Fisher-Yates with a 32-bit LCG seeded by the surface (host side, both builds)
inline void vxs_permutation(uint32_t seed, unsigned char perm[512])
{
	unsigned char p[256];
	for (int i = 0; i < 256; ++i) p[i] = (unsigned char)i;
	uint32_t s = seed * 2654435761u + 12345u;
	for (int i = 255; i > 0; --i)
	{
		s = s * 1664525u + 1013904223u;
		const int j = (int)((s >> 8) % (uint32_t)(i + 1));
		const unsigned char t = p[i]; p[i] = p[j]; p[j] = t;
	}
	for (int i = 0; i < 512; ++i) perm[i] = p[i & 255];
}

VXS_FN float vxs_fade(float t) { return t * t * t * (t * (t * 6.f - 15.f) + 10.f); }

// gradient of hash h dotted with (x, y, z): the 12 edge directions of the cube (4 of them twice)
VXS_FN float vxs_grad(int h, float x, float y, float z)
{
	const int k = h & 15;
	const float u = k < 8 ? x : y;
	const float v = k < 4 ? y : ((k == 12 || k == 14) ? x : z);
	return ((k & 1) ? -u : u) + ((k & 2) ? -v : v);
}

// classic gradient noise; perm = 512-entry table (vxs_permutation)
VXS_FN float vxs_noise3(const unsigned char* perm, float x, float y, float z)
{
	const float xf = VXS_FLOOR(x), yf = VXS_FLOOR(y), zf = VXS_FLOOR(z);
	const int xi = (int)xf & 255, yi = (int)yf & 255, zi = (int)zf & 255;
	const float dx = x - xf, dy = y - yf, dz = z - zf;
	const float u = vxs_fade(dx), v = vxs_fade(dy), w = vxs_fade(dz);
	float out = 0.f;
	for (int cz = 0; cz < 2; ++cz) for (int cy = 0; cy < 2; ++cy) for (int cx = 0; cx < 2; ++cx)
	{
		const int h = perm[perm[perm[xi + cx] + yi + cy] + zi + cz];
		const float dot = vxs_grad(h, dx - (float)cx, dy - (float)cy, dz - (float)cz);
		const float wgt = ((cx ? u : 1.f - u) * (cy ? v : 1.f - v)) * (cz ? w : 1.f - w);
		out = out + dot * wgt;
	}
	return out;
}

VXS_FN float vxs_noise2(const unsigned char* perm, float x, float y) { return vxs_noise3(perm, x, y, 0.5f); }

VXS_FN float vxs_fbm2(const unsigned char* perm, float x, float y)
{
	float total = 0.f, amp = 1.f, freq = 1.f;
	for (int o = 0; o < 5; ++o) // 5 octaves, lacunarity 2, gain 0.5
	{
		total = total + amp * vxs_noise2(perm, x * freq, y * freq);
		amp = amp * 0.5f;
		freq = freq * 2.f;
	}
	return total;
}

VXS_FN float vxs_clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

// Distance (linear, in voxels) + material + blend of surface `s` at (x, y, z), grid (Z-up) coordinates.
// perm is only read by VXB_SURFACE_TERRAIN.
VXS_FN float vxs_surface_value(const vxb_surface& s, const unsigned char* perm, float x, float y, float z, unsigned& mat, unsigned& blend)
{
	mat = s.material; blend = s.blend;
	if (s.kind == VXB_SURFACE_SPHERE)
	{
		// d = |p - c| - r, clamped to +-100 (keeps the reference's char conversion defined, VoxelGrid.cpp:37-40)
		const float dx = x - s.p[0], dy = y - s.p[1], dz = z - s.p[2];
		return vxs_clampf(VXS_SQRT((dx * dx + dy * dy) + dz * dz) - s.p[3], -100.f, 100.f);
	}
	if (s.kind == VXB_SURFACE_PLANE)
	{
		// d = n . p - d0 (n need not be unit length: the client decides the metric)
		return vxs_clampf(((s.p[0] * x + s.p[1] * y) + s.p[2] * z) - s.p[3], -100.f, 100.f);
	}
	// VXB_SURFACE_TERRAIN: heightfield from 5-octave fbm + 3-D detail noise; three material bands with
	// noisy boundaries, ore pockets, blends rising smoothly across the middle band.  p[0] = N, the terrain's horizontal period
	// (the grid edge it was designed for); p[1], p[2] = (x, y) origin of the window.
	const float N = s.p[0];
	const float gx = x + s.p[1], gy = y + s.p[2];
	const float height = 0.5f * N + (0.18f * N) * vxs_fbm2(perm, VXS_DIV(gx, N) * 4.f, VXS_DIV(gy, N) * 4.f);
	const float d = (z - height) + 6.f * vxs_noise3(perm, VXS_DIV(gx, 24.f), VXS_DIV(gy, 24.f), VXS_DIV(z, 24.f));
	const float h1 = 0.45f * N + (0.03f * N) * vxs_noise2(perm, VXS_DIV(gx, 37.f), VXS_DIV(gy, 37.f));
	const float h2 = 0.60f * N + (0.03f * N) * vxs_noise2(perm, VXS_DIV(gx, 53.f) + 7.7f, VXS_DIV(gy, 53.f) + 3.3f);
	unsigned m = z < h1 ? 0u : (z < h2 ? 1u : 2u);
	if (vxs_noise3(perm, VXS_DIV(gx, 48.f) + 11.1f, VXS_DIV(gy, 48.f) + 5.5f, VXS_DIV(z, 48.f) + 2.2f) > 0.35f) m = 3u;
	float span = h2 - h1;
	if (span < 1.f) span = 1.f;
	const float t = vxs_clampf(VXS_DIV(z - h1, span), 0.f, 1.f);
	mat = m;
	blend = (unsigned)(int)(255.f * (t * t * (3.f - 2.f * t)));
	return vxs_clampf(d, -100.f, 100.f);
}

// the reference's `round` + toGridDistValue (VoxelGrid.cpp:37-50): away from zero, then clamp; bound = 4 for a fill
// (Grid::Create), 127/-128 for an edit (InjectSurface does not apply the +-4 clamp, :441-452)
VXS_FN int vxs_round_away(float value)
{
#if defined(__CUDA_ARCH__)
	float a = ceilf(fabsf(value));
#else
	float a = std::ceil(std::fabs(value));
#endif
	if (a < -128.f) a = -128.f; // max_value((float)CHAR_MIN, ceil|v|): never binds, kept for fidelity
	float r = a * (value > 0.f ? 1.f : -1.f);
	if (r > 127.f) r = 127.f;
	return (int)r; // char(...) of a value in [-128, 127]
}
