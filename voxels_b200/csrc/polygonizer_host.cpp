// libvoxels_b200.so - drop-in backend for the reference's Polygonizer API.
//
// Defines exactly the symbols the reference's src/TransVoxelImpl.cpp defines for the public headers
// (compiled against /root/reference/include by -I, nothing copied):
//     Voxels::Polygonizer::Polygonizer / ~Polygonizer / Execute        include/Polygonizer.h:215-239  (impl TransVoxelImpl.cpp:64-79)
//     Voxels::Modification::Create / ~Modification                     include/Polygonizer.h:182-211  (impl :81-89)
//     Voxels::PolygonSurface::INVALID_ID                               include/Polygonizer.h:177      (impl :91)
//     Voxels::GetBlockExtent()                                         (impl :2171-2173)
// and returns objects implementing the pure-virtual interfaces PolygonSurface / BlockPolygons / Modification
// (same vtable order - they derive from the reference's own declarations).
//
// All computation happens on the GPU behind the C ABI of include/vxb200.h; this file only moves bytes:
//   Grid --(public accessors Grid::GetBlockDistanceData / GetBlockMaterialData, include/Grid.h:127-139)-->
//   pinned staging --> vxb_grid_upload_blocks --> vxb_polygonize --> vxb_result_download --> block views.
// There is no CPU polygonization path: if the device or the CUDA library is unavailable Execute logs an
// LS_Error and returns nullptr (the reference's own failure convention, TransVoxelImpl.cpp:2155-2164).
#include "stdafx.h" // reference src/stdafx.h through the shim: Logger (VOXLOG), VOXELS_LOG_SIZE
#include <Structs.h>
#include <Grid.h>
#include <MaterialMap.h>
#include <Polygonizer.h>

#include <cstdint>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/vxb200.h"

namespace Voxels
{

namespace
{
void logMessage(LogSeverity severity, const char* text)
{
	if (Logger::Get()) Logger::Get()->Log(severity, text);
}

struct HostBuffer // pinned, grows, never shrinks
{
	void* p = nullptr;
	size_t bytes = 0;
	bool ensure(size_t want)
	{
		if (want <= bytes) return true;
		if (p) vxb_host_free(p);
		p = vxb_host_alloc(want);
		bytes = p ? want : 0;
		return p != nullptr;
	}
	~HostBuffer() { if (p) vxb_host_free(p); }
};

struct BlockView : public BlockPolygons
{
	unsigned Id;
	const PolygonVertex* Vertices; unsigned VertexCount;
	const unsigned* Indices; unsigned IndexCount;
	const PolygonVertex* TransVertices[6]; unsigned TransVertexCount[6];
	const unsigned* TransIndices[6]; unsigned TransIndexCount[6];
	float3 MinimalCorner, MaximalCorner;

	virtual unsigned GetId() const override { return Id; }
	virtual const PolygonVertex* GetVertices(unsigned* count) const override { if (count) *count = VertexCount; return VertexCount ? Vertices : nullptr; }
	virtual const unsigned* GetIndices(unsigned* count) const override { if (count) *count = IndexCount; return IndexCount ? Indices : nullptr; }
	virtual const PolygonVertex* GetTransitionVertices(TransitionFaceId face, unsigned* count) const override
	{ if (count) *count = TransVertexCount[face]; return TransVertexCount[face] ? TransVertices[face] : nullptr; }
	virtual const unsigned* GetTransitionIndices(TransitionFaceId face, unsigned* count) const override
	{ if (count) *count = TransIndexCount[face]; return TransIndexCount[face] ? TransIndices[face] : nullptr; }
	virtual float3 GetMinimalCorner() const override { return MinimalCorner; }
	virtual float3 GetMaximalCorner() const override { return MaximalCorner; }
};

// The polygonized surface: block views over four host arenas filled by one vxb_result_download.
struct SurfaceImpl : public PolygonSurface
{
	float3 Extents;
	std::vector<std::vector<BlockView>> Levels;
	PolygonizationStatistics Stats;
	unsigned CacheBytes = 0;
	HostBuffer Verts, Idx, TransVerts, TransIdx;

	virtual float3 GetExtents() const override { return Extents; }
	virtual unsigned GetLevelsCount() const override { return unsigned(Levels.size()); }
	virtual unsigned GetBlocksForLevelCount(unsigned level) const override { return unsigned(Levels[level].size()); }
	virtual const BlockPolygons* GetBlockForLevel(unsigned level, unsigned id) const override
	{ return id < Levels[level].size() ? &Levels[level][id] : nullptr; }
	virtual const PolygonizationStatistics* GetStatistics() const override { return &Stats; }
	virtual unsigned GetCacheSizeBytes() const override { return CacheBytes; }
	virtual unsigned GetPolygonDataSizeBytes() const override
	{
		// mirrors PolygonMap::GetPolygonDataSizeBytes (TransVoxelImpl.cpp:222-235), including its habit of adding
		// the size of the six per-face vector objects instead of their contents
		size_t result = 0;
		for (const auto& level : Levels) for (const BlockView& b : level)
			result += size_t(b.VertexCount) * sizeof(PolygonVertex) + size_t(b.IndexCount) * sizeof(unsigned)
				+ 6 * sizeof(std::vector<PolygonVertex>) + 6 * sizeof(std::vector<unsigned>);
		return unsigned(result);
	}
	virtual void Destroy() override { delete this; }
};

struct ModificationImpl : public Modification
{
	std::vector<unsigned> ModifiedBlocks;
	virtual const unsigned* GetModifiedBlocks(unsigned* count) const override
	{ if (count) *count = unsigned(ModifiedBlocks.size()); return ModifiedBlocks.empty() ? nullptr : ModifiedBlocks.data(); }
	virtual void Destroy() override { delete this; }
};
}

// The opaque implementation class Polygonizer holds (include/Polygonizer.h:238).
class TransVoxelImpl
{
public:
	TransVoxelImpl() : Context(nullptr), Failed(false) {}
	~TransVoxelImpl() { if (Context) vxb_destroy(Context); }

	PolygonSurface* Execute(const Grid& grid, const MaterialMap* materials, Modification* modification);

private:
	bool EnsureContext();
	vxb_context* Context;
	bool Failed;
	HostBuffer StageDist, StageMat, StageBlend;
};

bool TransVoxelImpl::EnsureContext()
{
	if (Context) return true;
	if (Failed) return false;
	const int rc = vxb_create(0, &Context);
	if (rc != VXB_OK)
	{
		char buffer[VOXELS_LOG_SIZE];
		snprintf(buffer, VOXELS_LOG_SIZE, "Unable to polygonize grid: the B200 backend could not be initialised (%s)", vxb_last_error(nullptr));
		logMessage(LS_Error, buffer);
		Context = nullptr; Failed = true;
		return false;
	}
	return true;
}

PolygonSurface* TransVoxelImpl::Execute(const Grid& grid, const MaterialMap* materials, Modification* modification)
{
	char buffer[VOXELS_LOG_SIZE];
	const unsigned n = grid.GetWidth();
	if (grid.GetDepth() != n || grid.GetHeight() != n || n < 16 || (n & (n - 1)) != 0)
	{
		snprintf(buffer, VOXELS_LOG_SIZE, "Unable to polygonize grid: the grid must be a cube with a power-of-two edge >= 16 (got %u x %u x %u)", n, grid.GetDepth(), grid.GetHeight());
		logMessage(LS_Error, buffer);
		return nullptr;
	}
	if (!EnsureContext()) return nullptr;
	auto fail = [&](const char* what) -> PolygonSurface* {
		snprintf(buffer, VOXELS_LOG_SIZE, "Unable to polygonize grid: %s (%s)", what, vxb_last_error(Context));
		logMessage(LS_Error, buffer);
		return nullptr;
	};

	// ---- MaterialMap::GetMaterial pre-tabulated (the reference calls it per output vertex, :1249) ----
	uint8_t table[256 * 6], valid[256];
	for (unsigned id = 0; id < 256; ++id)
	{
		const MaterialMap::Material* m = materials ? materials->GetMaterial((unsigned char)id) : nullptr;
		valid[id] = m ? 1 : 0;
		for (int k = 0; k < 3; ++k) { table[id * 6 + k] = m ? m->DiffuseIds0[k] : 0; table[id * 6 + 3 + k] = m ? m->DiffuseIds1[k] : 0; }
	}
	if (vxb_set_materials(Context, table, valid) != VXB_OK) return fail("material table upload failed");

	// ---- grid -> device: every block through the public accessors, decompressed in parallel into pinned staging ----
	const unsigned nb = n / 16;
	const size_t volume = size_t(n) * n * n;
	if (!StageDist.ensure(volume) || !StageMat.ensure(volume) || !StageBlend.ensure(volume)) return fail("pinned staging allocation failed");
	{
		char* sd = static_cast<char*>(StageDist.p);
		unsigned char* sm = static_cast<unsigned char*>(StageMat.p);
		unsigned char* sb = static_cast<unsigned char*>(StageBlend.p);
		const long total = long(nb) * nb * nb;
		#pragma omp parallel for schedule(static)
		for (long b = 0; b < total; ++b)
		{
			const float3 coords(float(b % nb), float((b / nb) % nb), float(b / (long(nb) * nb)));
			grid.GetBlockDistanceData(coords, sd + size_t(b) * 4096);
			grid.GetBlockMaterialData(coords, sm + size_t(b) * 4096, sb + size_t(b) * 4096);
		}
	}
	if (vxb_grid_upload_blocks(Context, n, static_cast<const int8_t*>(StageDist.p), static_cast<const uint8_t*>(StageMat.p),
		static_cast<const uint8_t*>(StageBlend.p)) != VXB_OK) return fail("grid upload failed");

	// ---- polygonize on the device ----
	// Incremental updates (a Modification) are served by a full re-polygonization for now: the result is a valid
	// surface for the edited grid, but block ids restart and the reference's stale-cache quirks are not reproduced.
	if (vxb_polygonize(Context, 0, 0) != VXB_OK) return fail("polygonization failed");
	vxb_result_info info;
	if (vxb_result_info_get(Context, &info) != VXB_OK) return fail("no result");

	SurfaceImpl* surface = new SurfaceImpl;
	if (!surface->Verts.ensure(size_t(info.vertex_span) * sizeof(PolygonVertex) + 16) || !surface->Idx.ensure(size_t(info.index_span) * 4 + 16)
		|| !surface->TransVerts.ensure(size_t(info.trans_vertex_span) * sizeof(PolygonVertex) + 16) || !surface->TransIdx.ensure(size_t(info.trans_index_span) * 4 + 16))
	{ delete surface; return fail("host arena allocation failed"); }
	std::vector<vxb_block_record> records(info.block_count);
	if (vxb_result_download(Context, records.data(), surface->Verts.p, static_cast<uint32_t*>(surface->Idx.p), surface->TransVerts.p,
		static_cast<uint32_t*>(surface->TransIdx.p)) != VXB_OK)
	{ delete surface; return fail("result download failed"); }

	// ---- views ----
	surface->Extents = float3(float(n), float(n), float(n)); // (W, H, D) :481
	surface->Levels.resize(info.levels_total);
	const PolygonVertex* verts = static_cast<const PolygonVertex*>(surface->Verts.p);
	const unsigned* idx = static_cast<const unsigned*>(surface->Idx.p);
	const PolygonVertex* tverts = static_cast<const PolygonVertex*>(surface->TransVerts.p);
	const unsigned* tidx = static_cast<const unsigned*>(surface->TransIdx.p);
	for (const vxb_block_record& r : records)
	{
		BlockView b;
		b.Id = r.id;
		b.Vertices = verts + r.vertex_offset; b.VertexCount = r.vertex_count;
		b.Indices = idx + r.index_offset; b.IndexCount = r.index_count;
		for (int f = 0; f < 6; ++f)
		{
			b.TransVertices[f] = tverts + r.trans_vertex_offset[f]; b.TransVertexCount[f] = r.trans_vertex_count[f];
			b.TransIndices[f] = tidx + r.trans_index_offset[f]; b.TransIndexCount[f] = r.trans_index_count[f];
		}
		const unsigned m = 16u << r.level, nbl = n / m;
		const unsigned bx = r.coord_id % nbl, by = (r.coord_id / nbl) % nbl, bz = r.coord_id / (nbl * nbl);
		b.MinimalCorner = float3(float(bx * m), float(bz * m), float(by * m)); // y/z swapped on output (:1289-1291)
		b.MaximalCorner = float3(float(bx * m + m), float(bz * m + m), float(by * m + m));
		surface->Levels[r.level].push_back(b);
	}
	surface->Stats.BlocksCalculated = info.stats[0];
	surface->Stats.TrivialCells = info.stats[1];
	surface->Stats.NonTrivialCells = info.stats[2];
	surface->Stats.DegenerateTrianglesRemoved = info.stats[3];
	for (unsigned i = 0; i < PolygonizationStatistics::CASES_COUNT; ++i) surface->Stats.PerCaseCellsCount[i] = info.stats[4 + i];
	{
		// PolygonMap::GetCacheSizeBytes (:196-220): consistency bits of every level-0 block + {id, blend} of every coarser cell
		size_t total = (size_t(nb) * nb * nb * 4096) >> 3;
		for (unsigned l = 1; l < info.levels_total; ++l) { const size_t c = nb >> l; total += c * c * c * 4096 * 2; }
		surface->CacheBytes = unsigned(total);
	}

	// one LS_Error per vertex whose material has no mapping, as the reference logs them (:1364-1368)
	const uint64_t unmapped = vxb_result_unmapped_materials(Context, nullptr, 0);
	if (unmapped)
	{
		std::vector<uint8_t> ids(unmapped);
		vxb_result_unmapped_materials(Context, ids.data(), unmapped);
		for (uint8_t id : ids)
		{
			snprintf(buffer, VOXELS_LOG_SIZE, "Unable to assign textures on vertex with material id %u", unsigned(id));
			logMessage(LS_Error, buffer);
		}
	}

	if (modification)
	{
		ModificationImpl* mod = static_cast<ModificationImpl*>(modification);
		for (const auto& level : surface->Levels) for (const BlockView& b : level) mod->ModifiedBlocks.push_back(b.Id);
		if (modification->Map) modification->Map->Destroy(); // the caller's old surface is replaced
		modification->Map = surface;
	}
	return surface;
}

///////// PUBLIC INTERFACE //////////////

Polygonizer::Polygonizer() : m_Impl(new TransVoxelImpl) {}
Polygonizer::~Polygonizer() { delete m_Impl; }

PolygonSurface* Polygonizer::Execute(const Grid& grid, const MaterialMap* materials, Modification* modification)
{
	return m_Impl->Execute(grid, materials, modification);
}

Modification* Modification::Create()
{
	ModificationImpl* result = new ModificationImpl;
	result->Map = nullptr;
	return result;
}

Modification::~Modification() {}

const unsigned PolygonSurface::INVALID_ID = 0xFFFFFFFF;

unsigned GetBlockExtent() { return 16; }

}
