// Test harness: a flat C ABI over the reference's PUBLIC C++ API (include/Voxels.h).
//
// This file is compiled twice from the same source:
//   * oracle/_ref/libvxh_ref.so   - linked with the UNMODIFIED reference polygonizer (the oracle),
//   * build/libvxh_b200.so        - linked with this repo's drop-in Polygonizer backend (the product),
// so the parity tests drive both through exactly the calls a client of the reference makes:
// Grid::Create / Grid::InjectSurface / Polygonizer::Execute / PolygonSurface accessors
// (reference include/Grid.h:29-161, include/Polygonizer.h:136-239).
//
// Nothing here is shipped; it is test infrastructure.
#include <cstddef>
#include <cstdlib>
#include <Voxels.h>

#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#ifdef VXH_WITH_INTERNAL_GRID
// oracle build only: lets the tests read VoxelGrid::IsBlockEmpty (src/VoxelGrid.h:55)
#include "stdafx.h"
#include "VoxelGrid.h"
#endif

#ifdef _OPENMP
#include <omp.h>
#endif

// the built-in procedural surfaces of the product's device-side grid store, evaluated here on the host with bit-identical
// floats and served to the UNMODIFIED reference grid store through its VoxelSurface callback
#include "../../voxels_b200/csrc/vxb_surfaces.h"

using namespace Voxels;

namespace
{
// Serves a dense float SDF + material + blend volume through the reference's VoxelSurface callback
// (include/VoxelSurface.h:35-40).  Values are written x-fastest, then y, then z, which is the order
// VoxelGrid's constructor consumes them in (src/VoxelGrid.cpp:116-126).
struct DenseSurface : public VoxelSurface
{
	unsigned N;
	const float* Dist;
	const unsigned char* Mat;
	const unsigned char* Blend;

	virtual void GetSurface(float xStart, float xEnd, float xStep,
		float yStart, float yEnd, float yStep,
		float zStart, float zEnd, float zStep,
		float* output, unsigned char* materialid, unsigned char* blend) override
	{
		size_t id = 0;
		for (float z = zStart; z < zEnd; z += zStep)
		for (float y = yStart; y < yEnd; y += yStep)
		for (float x = xStart; x < xEnd; x += xStep)
		{
			const size_t g = (size_t(z) * N + size_t(y)) * N + size_t(x);
			output[id] = Dist[g];
			if (materialid) materialid[id] = Mat ? Mat[g] : 0;
			if (blend) blend[id] = Blend ? Blend[g] : 0;
			++id;
		}
	}
};

// d = |p - c| - r, clamped to +-100 (keeps the reference's char conversion defined, VoxelGrid.cpp:37-40)
struct SphereSurface : public VoxelSurface
{
	float Cx, Cy, Cz, R;
	unsigned char Material, BlendValue;

	virtual void GetSurface(float xStart, float xEnd, float xStep,
		float yStart, float yEnd, float yStep,
		float zStart, float zEnd, float zStep,
		float* output, unsigned char* materialid, unsigned char* blend) override
	{
		size_t id = 0;
		for (float z = zStart; z < zEnd; z += zStep)
		for (float y = yStart; y < yEnd; y += yStep)
		for (float x = xStart; x < xEnd; x += xStep)
		{
			const float dx = x - Cx, dy = y - Cy, dz = z - Cz;
			float d = std::sqrt(dx * dx + dy * dy + dz * dz) - R;
			if (d > 100.f) d = 100.f;
			if (d < -100.f) d = -100.f;
			output[id] = d;
			if (materialid) materialid[id] = Material;
			if (blend) blend[id] = BlendValue;
			++id;
		}
	}
};

// vxb_surface (include/vxb200.h) as a client VoxelSurface: values x-fastest, then y, then z (VoxelGrid.cpp:116-126)
struct BuiltinSurface : public VoxelSurface
{
	vxb_surface S;
	unsigned char Perm[512];
	explicit BuiltinSurface(const vxb_surface& s) : S(s) { vxs_permutation(s.seed, Perm); }

	virtual void GetSurface(float xStart, float xEnd, float xStep,
		float yStart, float yEnd, float yStep,
		float zStart, float zEnd, float zStep,
		float* output, unsigned char* materialid, unsigned char* blend) override
	{
		size_t id = 0;
		// sample k of an axis sits at start + k * step (the grid hands out whole blocks / whole sections: VoxelGrid.cpp:100-113, :419-428)
		int nx = 0, ny = 0, nz = 0;
		for (float x = xStart; x < xEnd; x += xStep) ++nx;
		for (float y = yStart; y < yEnd; y += yStep) ++ny;
		for (float z = zStart; z < zEnd; z += zStep) ++nz;
		for (int kz = 0; kz < nz; ++kz)
		for (int ky = 0; ky < ny; ++ky)
		for (int kx = 0; kx < nx; ++kx)
		{
			unsigned m, b;
			output[id] = vxs_surface_value(S, Perm, xStart + float(kx) * xStep, yStart + float(ky) * yStep, zStart + float(kz) * zStep, m, b);
			if (materialid) materialid[id] = (unsigned char)m;
			if (blend) blend[id] = (unsigned char)b;
			++id;
		}
	}
};

struct TableMaterialMap : public MaterialMap
{
	Material Table[256];
	bool Valid[256];
	virtual Material* GetMaterial(unsigned char id) const override
	{
		return Valid[id] ? const_cast<Material*>(&Table[id]) : nullptr;
	}
};

struct BlockRow // 21 x 4 bytes, mirrored by tests/harness.py
{
	uint32_t Id;
	float Min[3];
	float Max[3];
	uint32_t VertexCount;
	uint32_t IndexCount;
	uint32_t TransVertexCount[6];
	uint32_t TransIndexCount[6];
};

bool g_Initialized = false;
unsigned g_ErrorLogs = 0;
void LogSink(LogSeverity severity, const char* message)
{
	if (severity >= LS_Error) {
		++g_ErrorLogs;
		if (g_ErrorLogs < 8) fprintf(stderr, "[voxels] %s\n", message);
	}
}
}

extern "C"
{

int vxh_init()
{
	if (g_Initialized) return 0;
	const auto err = InitializeVoxels(VOXELS_VERSION, &LogSink, nullptr);
	g_Initialized = (err == IE_Ok);
	return int(err);
}

unsigned vxh_error_logs() { return g_ErrorLogs; }

void vxh_set_threads(int n)
{
#ifdef _OPENMP
	if (n > 0) omp_set_num_threads(n);
#else
	(void)n;
#endif
}

int vxh_max_threads()
{
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}

// ---- grids -------------------------------------------------------------------------------------

// Reference-quantised grid from a float SDF (round + clamp to +-4 happens inside the reference).
void* vxh_grid_create_from_floats(unsigned n, const float* dist, const unsigned char* mat, const unsigned char* blend)
{
	DenseSurface s;
	s.N = n; s.Dist = dist; s.Mat = mat; s.Blend = blend;
	return Grid::Create(n, n, n, 0.f, 0.f, 0.f, 1.f, &s);
}

void* vxh_grid_create_sphere(unsigned n, float cx, float cy, float cz, float r, unsigned char material, unsigned char blend)
{
	SphereSurface s;
	s.Cx = cx; s.Cy = cy; s.Cz = cz; s.R = r; s.Material = material; s.BlendValue = blend;
	return Grid::Create(n, n, n, 0.f, 0.f, 0.f, 1.f, &s);
}

// Grid::Create(n, n, n, sx, sy, sz, step, &builtin): the reference's own constructor walks the blocks and quantises
void* vxh_grid_create_builtin(unsigned n, const vxb_surface* surface, float sx, float sy, float sz, float step)
{
	BuiltinSurface s(*surface);
	return Grid::Create(n, n, n, sx, sy, sz, step, &s);
}

// The same voxels as dense arrays, on every host thread (for full-size inputs: the reference's constructor is serial):
// surface value -> the reference's round (away from zero) -> clamp to +-4 (VoxelGrid.cpp:37-50), start 0, step 1.
void vxh_builtin_dense(unsigned n, const vxb_surface* surface, signed char* dist, unsigned char* mat, unsigned char* blend)
{
	BuiltinSurface s(*surface);
	#pragma omp parallel for schedule(static)
	for (long zy = 0; zy < long(n) * n; ++zy)
	{
		const unsigned z = unsigned(zy / n), y = unsigned(zy % n);
		for (unsigned x = 0; x < n; ++x)
		{
			unsigned m, b;
			const float d = vxs_surface_value(s.S, s.Perm, float(x), float(y), float(z), m, b);
			int v = vxs_round_away(d);
			v = v > 4 ? 4 : (v < -4 ? -4 : v);
			const size_t g = (size_t(z) * n + y) * n + x;
			dist[g] = (signed char)v; mat[g] = (unsigned char)m; blend[g] = (unsigned char)b;
		}
	}
}

// Exact-bytes grid: empty grid + Grid::ModifyBlock*Data per block (include/Grid.h:135-146).
void* vxh_grid_from_dense(unsigned n, const signed char* dist, const unsigned char* mat, const unsigned char* blend)
{
	Grid* grid = Grid::Create(n, n, n);
	if (!grid) return nullptr;
	const unsigned nb = n / 16;
	char d[4096];
	unsigned char m[4096], b[4096];
	for (unsigned bz = 0; bz < nb; ++bz)
	for (unsigned by = 0; by < nb; ++by)
	for (unsigned bx = 0; bx < nb; ++bx)
	{
		for (unsigned z = 0; z < 16; ++z)
		for (unsigned y = 0; y < 16; ++y)
		{
			const size_t g = ((size_t(bz) * 16 + z) * n + (size_t(by) * 16 + y)) * n + size_t(bx) * 16;
			const unsigned l = z * 256 + y * 16;
			memcpy(d + l, dist + g, 16);
			if (mat) memcpy(m + l, mat + g, 16); else memset(m + l, 0, 16);
			if (blend) memcpy(b + l, blend + g, 16); else memset(b + l, 0, 16);
		}
		const float3 coords((float)bx, (float)by, (float)bz);
		grid->ModifyBlockDistanceData(coords, d);
		grid->ModifyBlockMaterialData(coords, m, b);
	}
	return grid;
}

unsigned vxh_grid_size(void* grid) { return static_cast<Grid*>(grid)->GetWidth(); }

void vxh_grid_to_dense(void* gridPtr, signed char* dist, unsigned char* mat, unsigned char* blend)
{
	Grid* grid = static_cast<Grid*>(gridPtr);
	const unsigned n = grid->GetWidth();
	const unsigned nb = n / 16;
	char d[4096];
	unsigned char m[4096], b[4096];
	for (unsigned bz = 0; bz < nb; ++bz)
	for (unsigned by = 0; by < nb; ++by)
	for (unsigned bx = 0; bx < nb; ++bx)
	{
		const float3 coords((float)bx, (float)by, (float)bz);
		grid->GetBlockDistanceData(coords, d);
		grid->GetBlockMaterialData(coords, m, b);
		for (unsigned z = 0; z < 16; ++z)
		for (unsigned y = 0; y < 16; ++y)
		{
			const size_t g = ((size_t(bz) * 16 + z) * n + (size_t(by) * 16 + y)) * n + size_t(bx) * 16;
			const unsigned l = z * 256 + y * 16;
			if (dist) memcpy(dist + g, d + l, 16);
			if (mat) memcpy(mat + g, m + l, 16);
			if (blend) memcpy(blend + g, b + l, 16);
		}
	}
}

// flags[blockId] = VoxelGrid::IsBlockEmpty; returns 0 when the private header is not compiled in.
int vxh_grid_empty_flags(void* gridPtr, unsigned char* flags)
{
#ifdef VXH_WITH_INTERNAL_GRID
	Grid* grid = static_cast<Grid*>(gridPtr);
	const unsigned nb = grid->GetWidth() / 16;
	VoxelGrid* internal = grid->GetInternalRepresentation();
	for (unsigned bz = 0; bz < nb; ++bz)
	for (unsigned by = 0; by < nb; ++by)
	for (unsigned bx = 0; bx < nb; ++bx)
		flags[(bz * nb + by) * nb + bx] = internal->IsBlockEmpty(glm::vec3(bx, by, bz)) ? 1 : 0;
	return 1;
#else
	(void)gridPtr; (void)flags;
	return 0;
#endif
}

// Sphere edit through Grid::InjectSurface; out6 = returned box (already y/z swizzled by the reference).
void vxh_grid_inject_sphere(void* gridPtr, float px, float py, float pz, float radius, float extent, int type, float* out6)
{
	Grid* grid = static_cast<Grid*>(gridPtr);
	SphereSurface s;
	s.Cx = 0.f; s.Cy = 0.f; s.Cz = 0.f; s.R = radius; s.Material = 0; s.BlendValue = 0;
	const float3pair box = grid->InjectSurface(float3(px, py, pz), float3(extent, extent, extent), &s, InjectionType(type));
	out6[0] = box.first.x; out6[1] = box.first.y; out6[2] = box.first.z;
	out6[3] = box.second.x; out6[4] = box.second.y; out6[5] = box.second.z;
}

// Grid::InjectSurface with a built-in surface (sampled relative to the position, VoxelGrid.cpp:419-428)
void vxh_grid_inject_builtin(void* gridPtr, const float* pos, const float* ext, const vxb_surface* surface, int type, float* out6)
{
	Grid* grid = static_cast<Grid*>(gridPtr);
	BuiltinSurface s(*surface);
	const float3pair box = grid->InjectSurface(float3(pos[0], pos[1], pos[2]), float3(ext[0], ext[1], ext[2]), &s, InjectionType(type));
	out6[0] = box.first.x; out6[1] = box.first.y; out6[2] = box.first.z;
	out6[3] = box.second.x; out6[4] = box.second.y; out6[5] = box.second.z;
}

void vxh_grid_inject_material(void* gridPtr, const float* pos, const float* ext, unsigned material, int addSubtractBlend, float* out6)
{
	Grid* grid = static_cast<Grid*>(gridPtr);
	const float3pair box = grid->InjectMaterial(float3(pos[0], pos[1], pos[2]), float3(ext[0], ext[1], ext[2]), (unsigned char)material, addSubtractBlend != 0);
	out6[0] = box.first.x; out6[1] = box.first.y; out6[2] = box.first.z;
	out6[3] = box.second.x; out6[4] = box.second.y; out6[5] = box.second.z;
}

unsigned vxh_grid_pack_size(void* gridPtr, void** packOut)
{
	Grid::PackedGrid* pack = static_cast<Grid*>(gridPtr)->PackForSave();
	*packOut = pack;
	return pack->GetSize();
}
const char* vxh_pack_data(void* pack) { return static_cast<Grid::PackedGrid*>(pack)->GetData(); }
void vxh_pack_destroy(void* pack) { static_cast<Grid::PackedGrid*>(pack)->Destroy(); }
void* vxh_grid_load(const char* blob, unsigned size) { return Grid::Load(blob, size); }

void vxh_grid_destroy(void* grid) { static_cast<Grid*>(grid)->Destroy(); }

// ---- polygonization ----------------------------------------------------------------------------

void* vxh_modification_create() { return Modification::Create(); }
void vxh_modification_destroy(void* m) { static_cast<Modification*>(m)->Destroy(); }
unsigned vxh_modification_blocks(void* m, unsigned* out, unsigned capacity)
{
	unsigned count = 0;
	const unsigned* ids = static_cast<Modification*>(m)->GetModifiedBlocks(&count);
	for (unsigned i = 0; i < count && i < capacity; ++i) out[i] = ids[i];
	return count;
}

// materialTable: 256 x 6 bytes (DiffuseIds0[3], DiffuseIds1[3]) or null for the identity map;
// validMask: 256 bytes or null (all valid).  mod: null for a full run, else box6 + surface are set on it.
void* vxh_polygonize(void* gridPtr, const unsigned char* materialTable, const unsigned char* validMask,
	void* modPtr, void* surfaceForMod, const float* box6, double* seconds)
{
	Grid* grid = static_cast<Grid*>(gridPtr);
	TableMaterialMap map;
	for (unsigned i = 0; i < 256; ++i)
	{
		for (unsigned k = 0; k < 3; ++k)
		{
			map.Table[i].DiffuseIds0[k] = materialTable ? materialTable[i * 6 + k] : (unsigned char)i;
			map.Table[i].DiffuseIds1[k] = materialTable ? materialTable[i * 6 + 3 + k] : (unsigned char)i;
		}
		map.Valid[i] = validMask ? validMask[i] != 0 : true;
	}
	Modification* mod = static_cast<Modification*>(modPtr);
	if (mod)
	{
		mod->Map = static_cast<PolygonSurface*>(surfaceForMod);
		mod->MinCornerModified = float3(box6[0], box6[1], box6[2]);
		mod->MaxCornerModified = float3(box6[3], box6[4], box6[5]);
	}
	Polygonizer polygonizer;
	const auto t0 = std::chrono::steady_clock::now();
	PolygonSurface* surface = polygonizer.Execute(*grid, &map, mod);
	const auto t1 = std::chrono::steady_clock::now();
	if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
	return surface;
}

void vxh_surface_destroy(void* s) { static_cast<PolygonSurface*>(s)->Destroy(); }
unsigned vxh_surface_levels(void* s) { return static_cast<PolygonSurface*>(s)->GetLevelsCount(); }
unsigned vxh_surface_blocks(void* s, unsigned level) { return static_cast<PolygonSurface*>(s)->GetBlocksForLevelCount(level); }
void vxh_surface_extents(void* s, float* out3)
{
	const float3 e = static_cast<PolygonSurface*>(s)->GetExtents();
	out3[0] = e.x; out3[1] = e.y; out3[2] = e.z;
}
unsigned vxh_surface_cache_bytes(void* s) { return static_cast<PolygonSurface*>(s)->GetCacheSizeBytes(); }
unsigned vxh_surface_polygon_bytes(void* s) { return static_cast<PolygonSurface*>(s)->GetPolygonDataSizeBytes(); }

// out: BlocksCalculated, TrivialCells, NonTrivialCells, DegenerateTrianglesRemoved, PerCaseCellsCount[16]
void vxh_surface_stats(void* s, unsigned* out20)
{
	const PolygonizationStatistics* st = static_cast<PolygonSurface*>(s)->GetStatistics();
	out20[0] = st->BlocksCalculated;
	out20[1] = st->TrivialCells;
	out20[2] = st->NonTrivialCells;
	out20[3] = st->DegenerateTrianglesRemoved;
	for (unsigned i = 0; i < 16; ++i) out20[4 + i] = st->PerCaseCellsCount[i];
}

// totals4: vertices, indices, transition vertices, transition indices over all blocks of the level
void vxh_surface_level_totals(void* s, unsigned level, uint64_t* totals4)
{
	PolygonSurface* surface = static_cast<PolygonSurface*>(s);
	totals4[0] = totals4[1] = totals4[2] = totals4[3] = 0;
	const unsigned count = surface->GetBlocksForLevelCount(level);
	for (unsigned i = 0; i < count; ++i)
	{
		const BlockPolygons* block = surface->GetBlockForLevel(level, i);
		unsigned c = 0;
		block->GetVertices(&c); totals4[0] += c;
		block->GetIndices(&c); totals4[1] += c;
		for (int f = 0; f < 6; ++f)
		{
			block->GetTransitionVertices(BlockPolygons::TransitionFaceId(f), &c); totals4[2] += c;
			block->GetTransitionIndices(BlockPolygons::TransitionFaceId(f), &c); totals4[3] += c;
		}
	}
}

// Flattens one LOD level: block table rows + concatenated arrays (block order, then face order).
void vxh_surface_level_dump(void* s, unsigned level, void* rowsOut,
	void* vertices, unsigned* indices, void* transVertices, unsigned* transIndices)
{
	PolygonSurface* surface = static_cast<PolygonSurface*>(s);
	BlockRow* rows = static_cast<BlockRow*>(rowsOut);
	PolygonVertex* v = static_cast<PolygonVertex*>(vertices);
	PolygonVertex* tv = static_cast<PolygonVertex*>(transVertices);
	const unsigned count = surface->GetBlocksForLevelCount(level);
	for (unsigned i = 0; i < count; ++i)
	{
		const BlockPolygons* block = surface->GetBlockForLevel(level, i);
		BlockRow& row = rows[i];
		row.Id = block->GetId();
		const float3 mn = block->GetMinimalCorner(), mx = block->GetMaximalCorner();
		row.Min[0] = mn.x; row.Min[1] = mn.y; row.Min[2] = mn.z;
		row.Max[0] = mx.x; row.Max[1] = mx.y; row.Max[2] = mx.z;
		unsigned c = 0;
		const PolygonVertex* pv = block->GetVertices(&c);
		row.VertexCount = c;
		if (c) { memcpy(v, pv, size_t(c) * sizeof(PolygonVertex)); v += c; }
		const unsigned* pi = block->GetIndices(&c);
		row.IndexCount = c;
		if (c) { memcpy(indices, pi, size_t(c) * 4); indices += c; }
		for (int f = 0; f < 6; ++f)
		{
			const PolygonVertex* ptv = block->GetTransitionVertices(BlockPolygons::TransitionFaceId(f), &c);
			row.TransVertexCount[f] = c;
			if (c) { memcpy(tv, ptv, size_t(c) * sizeof(PolygonVertex)); tv += c; }
			const unsigned* pti = block->GetTransitionIndices(BlockPolygons::TransitionFaceId(f), &c);
			row.TransIndexCount[f] = c;
			if (c) { memcpy(transIndices, pti, size_t(c) * 4); transIndices += c; }
		}
	}
}

} // extern "C"
