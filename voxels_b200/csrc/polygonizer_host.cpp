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
//   full run:        the grid store's compressed blocks --(parallel gather, PackForSave layout)--> pinned blob
//                    --> vxb_grid_upload_packed_streamed (copied and run-length decoded on the GPU while the next
//                    slab is being gathered) --> vxb_polygonize --> vxb_result_download --> block views;
//   incremental run: Grid::GetBlockDistanceData / GetBlockMaterialData (include/Grid.h:127-139) of the dirty blocks
//                    --> vxb_grid_update_blocks --> vxb_polygonize_region --> download --> splice.
// There is no CPU polygonization path: if the device or the CUDA library is unavailable Execute logs an
// LS_Error and returns nullptr (the reference's own failure convention, TransVoxelImpl.cpp:2155-2164).
#include "stdafx.h" // reference src/stdafx.h through the shim: Logger (VOXLOG), VOXELS_LOG_SIZE
#include <Structs.h>
#include <Grid.h>
// The grid store's compressed blocks (src/VoxelGrid.h:82-104: flags + three run-length coded channels per block) are what
// PackForSave serialises (src/VoxelGrid.cpp:269-315) and what the GPU decodes; the reference's own backend is a friend-less
// client of VoxelGrid too, but only needs decompressed blocks.  A maintainer adds `const std::vector<Block>& GetBlocks() const`
// (INTEGRATION.md); until then this translation unit reads the private member.
#define private public
#include <VoxelGrid.h>
#undef private
#include <MaterialMap.h>
#include <Polygonizer.h>

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <memory>
#include <mutex>
#include <new>
#include <vector>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <omp.h>

#include "../../include/vxb200.h"

namespace Voxels
{

namespace
{
void logMessage(LogSeverity severity, const char* text)
{
	if (Logger::Get()) Logger::Get()->Log(severity, text);
}

// Host memory that grows and never shrinks.  Large buffers are page-locked (full-speed DMA for the GB-sized uploads
// and downloads of a full run); small ones - the arenas of incremental runs, a few hundred KB per edit - are plain
// malloc, because pinning costs milliseconds per allocation.
struct HostBuffer
{
	void* p = nullptr;
	size_t bytes = 0;
	bool pinned = false;
	bool ensure(size_t want)
	{
		if (want <= bytes) return true;
		release();
		pinned = want >= (size_t(8) << 20);
		p = pinned ? vxb_host_alloc(want) : malloc(want ? want : 1);
		bytes = p ? want : 0;
		return p != nullptr;
	}
	void release() { if (p) { if (pinned) vxb_host_free(p); else free(p); } p = nullptr; bytes = 0; }
	~HostBuffer() { release(); }
};

struct Arena;

struct BlockView : public BlockPolygons
{
	unsigned Id;
	Arena* Owner; // the host arena this block's arrays live in
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

// Device contexts are expensive to create (stream, events, GBs of device memory), so destroyed surfaces hand theirs
// back to a process-wide pool.
std::mutex g_PoolLock;
std::vector<vxb_context*> g_ContextPool;

vxb_context* acquireContext()
{
	{
		std::lock_guard<std::mutex> guard(g_PoolLock);
		if (!g_ContextPool.empty()) { vxb_context* c = g_ContextPool.back(); g_ContextPool.pop_back(); return c; }
	}
	vxb_context* c = nullptr;
	return vxb_create(0, &c) == VXB_OK ? c : nullptr;
}

void releaseContext(vxb_context* c)
{
	if (!c) return;
	std::lock_guard<std::mutex> guard(g_PoolLock);
	if (g_ContextPool.size() < 2) g_ContextPool.push_back(c); else vxb_destroy(c);
}

// The page-locked staging of a full upload (3 n^3 bytes) costs tens of milliseconds to allocate, and clients create a
// Polygonizer per call: the buffers live in a process-wide pool like the contexts (never freed at exit on purpose: the
// CUDA runtime may already be gone when static destructors run).
struct Staging
{
	HostBuffer Dist, Mat, Blend; // decompressed blocks of the edited region (incremental runs)
	HostBuffer Blob;             // the grid in PackForSave form (full runs)
};
std::vector<Staging*> g_StagingPool;

Staging* acquireStaging()
{
	{
		std::lock_guard<std::mutex> guard(g_PoolLock);
		if (!g_StagingPool.empty()) { Staging* s = g_StagingPool.back(); g_StagingPool.pop_back(); return s; }
	}
	return new Staging;
}

void releaseStaging(Staging* s)
{
	if (!s) return;
	std::lock_guard<std::mutex> guard(g_PoolLock);
	if (g_StagingPool.size() < 2) g_StagingPool.push_back(s); else delete s;
}

// host copy of the output of ONE device run.  Live = block views still pointing into it: an incremental Execute erases
// the blocks of the dirty boxes (:443-450), and an arena whose last block is gone is freed (the reference frees the
// erased blocks' vectors the same way).
struct Arena
{
	HostBuffer Verts, Idx, TransVerts, TransIdx;
	size_t Live = 0;
};

// Page-locked arenas of full runs (hundreds of MB) are pooled like the contexts: allocating them costs ~0.1 s
std::vector<Arena*> g_ArenaPool;

Arena* acquireArena()
{
	{
		std::lock_guard<std::mutex> guard(g_PoolLock);
		if (!g_ArenaPool.empty()) { Arena* a = g_ArenaPool.back(); g_ArenaPool.pop_back(); a->Live = 0; return a; }
	}
	return new Arena;
}

void releaseArena(Arena* a)
{
	if (!a) return;
	if (a->Verts.pinned)
	{
		std::lock_guard<std::mutex> guard(g_PoolLock);
		if (g_ArenaPool.size() < 2) { g_ArenaPool.push_back(a); return; }
	}
	delete a;
}

struct ArenaReturn { void operator()(Arena* a) const { releaseArena(a); } };
typedef std::unique_ptr<Arena, ArenaReturn> ArenaPtr;

// The polygonized surface: block views over host arenas (one per run: the full run + one small one per edit).
// It owns the device context that holds the grid copy and the material caches of this surface, the way the
// reference's PolygonMap owns its MaterialCache (TransVoxelImpl.h:81-95), so later incremental runs find them.
struct SurfaceImpl final : public PolygonSurface
{
	float3 Extents;
	std::vector<std::vector<BlockView>> Levels;
	PolygonizationStatistics Stats;
	unsigned CacheBytes = 0;
	std::vector<ArenaPtr> Arenas;
	vxb_context* Context = nullptr;
	~SurfaceImpl() { releaseContext(Context); }

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

struct ModificationImpl final : public Modification
{
	std::vector<unsigned> ModifiedBlocks;
	virtual const unsigned* GetModifiedBlocks(unsigned* count) const override
	{ if (count) *count = unsigned(ModifiedBlocks.size()); return ModifiedBlocks.empty() ? nullptr : ModifiedBlocks.data(); }
	virtual void Destroy() override { delete this; }
};
}

namespace
{
// Threads for the gather of the compressed blocks (memory-bound: more than ~32 threads add nothing).  A container's CPU
// quota (cgroup v2 cpu.max) caps it further: a team larger than the quota burns the period's budget while spinning at
// the region's end and the whole process is then throttled for the rest of the period - measured on the B200 hosts as
// 60 ms stalls in whatever stage came next.  VXB200_PACK_THREADS overrides.
int packThreads()
{
	static const int threads = [] {
		if (const char* e = getenv("VXB200_PACK_THREADS")) { const int v = atoi(e); if (v > 0) return v; }
		int t = std::min(omp_get_max_threads(), 32);
		if (FILE* f = fopen("/sys/fs/cgroup/cpu.max", "r"))
		{
			long long quota = 0, period = 0;
			if (fscanf(f, "%lld %lld", &quota, &period) == 2 && quota > 0 && period > 0)
				t = std::min<long long>(t, std::max<long long>(1, quota / period));
			fclose(f);
		}
		return std::max(t, 1);
	}();
	return threads;
}

// Byte offset of every block in the PackForSave form (VoxelGrid.cpp:269-315: 16-byte header, 3 sizes per block, then per
// block {flags, distance, material, blend}); returns the total size.
size_t blobLayout(const std::vector<VoxelGrid::Block>& blocks, std::vector<size_t>& offsets)
{
	// the per-block sizes are read in parallel (a pass over the 80-byte Block records); the prefix sum is serial over 8 bytes
	// per block
	const size_t count = blocks.size();
	offsets.resize(count + 1);
	#pragma omp parallel for schedule(static) num_threads(packThreads())
	for (long b = 0; b < long(count); ++b)
		offsets[b] = size_t(4) + blocks[b].DistanceData.size() + blocks[b].MaterialData.size() + blocks[b].BlendData.size();
	size_t off = 16 + count * 12;
	for (size_t b = 0; b < count; ++b) { const size_t bytes = offsets[b]; offsets[b] = off; off += bytes; }
	offsets[count] = off;
	return off;
}

void packHeader(unsigned n, unsigned char* blob)
{
	const uint32_t header[4] = { 1u, n, n, n };
	memcpy(blob, header, 16);
}

// size-table entries + data of blocks [first, last)
void packBlocks(const std::vector<VoxelGrid::Block>& blocks, const std::vector<size_t>& offsets, size_t first, size_t last, unsigned char* blob)
{
	#pragma omp parallel for schedule(static) num_threads(packThreads())
	for (long b = long(first); b < long(last); ++b)
	{
		if (b + 8 < long(last))
		{
			// the channels are separate heap allocations: a dependent miss each unless asked for a few blocks ahead
			const auto& ahead = blocks[b + 8];
			__builtin_prefetch(ahead.DistanceData.data()); __builtin_prefetch(ahead.MaterialData.data()); __builtin_prefetch(ahead.BlendData.data());
		}
		const auto& blk = blocks[b];
		const uint32_t sz[3] = { uint32_t(blk.DistanceData.size()), uint32_t(blk.MaterialData.size()), uint32_t(blk.BlendData.size()) };
		memcpy(blob + 16 + size_t(b) * 12, sz, 12);
		unsigned char* out = blob + offsets[b];
		const uint32_t flags = blk.Flags;
		memcpy(out, &flags, 4); out += 4;
		memcpy(out, blk.DistanceData.data(), sz[0]); out += sz[0];
		memcpy(out, blk.MaterialData.data(), sz[1]); out += sz[1];
		memcpy(out, blk.BlendData.data(), sz[2]);
	}
}
}

// The gather on its own (tests, tools): the grid in PackForSave form into `out`; returns the size it takes (also when
// `out` is null or `capacity` too small: nothing is written then), 0 for a grid without block data.
extern "C" size_t voxels_b200_pack_grid(const Grid* grid, unsigned char* out, size_t capacity)
{
	if (!grid) return 0;
	const auto& blocks = grid->GetInternalRepresentation()->m_Blocks;
	if (blocks.empty()) return 0;
	std::vector<size_t> offsets;
	const size_t total = blobLayout(blocks, offsets);
	if (!out || capacity < total) return total;
	packHeader(grid->GetWidth(), out);
	packBlocks(blocks, offsets, 0, blocks.size(), out);
	return total;
}

// The opaque implementation class Polygonizer holds (include/Polygonizer.h:238).
class TransVoxelImpl
{
public:
	PolygonSurface* Execute(const Grid& grid, const MaterialMap* materials, Modification* modification);

private:
	PolygonSurface* ExecuteIncremental(const Grid& grid, SurfaceImpl* surface, ModificationImpl* modification);
	Staging* m_Staging = nullptr;

public:
	~TransVoxelImpl() { releaseStaging(m_Staging); }
};

namespace
{
// Downloads the context's current result into a new arena and builds the block views over it, per level, in the order
// PushBlocksToResult appends them (level, then z,y,x: :1274-1293).  Nothing of the surface is touched: a failure leaves it intact.
struct Downloaded
{
	ArenaPtr Storage;
	std::vector<std::vector<BlockView>> Levels;
};

bool downloadResult(vxb_context* ctx, unsigned n, const vxb_result_info& info, Downloaded& out)
{
	ArenaPtr arena(acquireArena());
	if (!arena->Verts.ensure(size_t(info.vertex_span) * sizeof(PolygonVertex) + 16) || !arena->Idx.ensure(size_t(info.index_span) * 4 + 16)
		|| !arena->TransVerts.ensure(size_t(info.trans_vertex_span) * sizeof(PolygonVertex) + 16) || !arena->TransIdx.ensure(size_t(info.trans_index_span) * 4 + 16))
		return false;
	std::vector<vxb_block_record> records(info.block_count);
	// the directory arrives first; the views are built while the arenas are still moving
	if (vxb_result_download_begin(ctx, records.data(), arena->Verts.p, static_cast<uint32_t*>(arena->Idx.p), arena->TransVerts.p,
		static_cast<uint32_t*>(arena->TransIdx.p)) != VXB_OK)
		return false;
	const PolygonVertex* verts = static_cast<const PolygonVertex*>(arena->Verts.p);
	const unsigned* idx = static_cast<const unsigned*>(arena->Idx.p);
	const PolygonVertex* tverts = static_cast<const PolygonVertex*>(arena->TransVerts.p);
	const unsigned* tidx = static_cast<const unsigned*>(arena->TransIdx.p);
	out.Levels.assign(info.levels_total, std::vector<BlockView>());
	{
		std::vector<size_t> perLevel(info.levels_total, 0);
		for (const vxb_block_record& r : records) if (r.level < info.levels_total) ++perLevel[r.level];
		for (unsigned l = 0; l < info.levels_total; ++l) out.Levels[l].reserve(perLevel[l]);
	}
	for (const vxb_block_record& r : records)
	{
		BlockView b;
		b.Id = r.id;
		b.Owner = arena.get();
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
		out.Levels[r.level].push_back(b);
	}
	if (vxb_result_download_end(ctx) != VXB_OK) return false;
	arena->Live = records.size();
	out.Storage = std::move(arena);
	return true;
}

// appends the downloaded blocks to the surface; arenas no block points into any more are freed
void commitResult(SurfaceImpl* surface, Downloaded& dl, const vxb_result_info& info)
{
	for (size_t l = 0; l < dl.Levels.size() && l < surface->Levels.size(); ++l)
		surface->Levels[l].insert(surface->Levels[l].end(), dl.Levels[l].begin(), dl.Levels[l].end());
	if (dl.Storage->Live) surface->Arenas.push_back(std::move(dl.Storage));
	surface->Arenas.erase(std::remove_if(surface->Arenas.begin(), surface->Arenas.end(),
		[](const ArenaPtr& a) { return a->Live == 0; }), surface->Arenas.end());
	surface->Stats.BlocksCalculated = info.stats[0];
	surface->Stats.TrivialCells = info.stats[1];
	surface->Stats.NonTrivialCells = info.stats[2];
	surface->Stats.DegenerateTrianglesRemoved = info.stats[3];
	for (unsigned i = 0; i < PolygonizationStatistics::CASES_COUNT; ++i) surface->Stats.PerCaseCellsCount[i] = info.stats[4 + i];
}

// one LS_Error per vertex whose material has no mapping, as the reference logs them (:1364-1368)
void logUnmapped(vxb_context* ctx)
{
	const uint64_t unmapped = vxb_result_unmapped_materials(ctx, nullptr, 0);
	if (!unmapped) return;
	std::vector<uint8_t> ids(unmapped);
	vxb_result_unmapped_materials(ctx, ids.data(), unmapped);
	char buffer[VOXELS_LOG_SIZE];
	for (uint8_t id : ids)
	{
		snprintf(buffer, VOXELS_LOG_SIZE, "Unable to assign textures on vertex with material id %u", unsigned(id));
		logMessage(LS_Error, buffer);
	}
}

bool uploadMaterials(vxb_context* ctx, const MaterialMap* materials)
{
	// MaterialMap::GetMaterial pre-tabulated (the reference calls it per output vertex, :1249)
	uint8_t table[256 * 6], valid[256];
	for (unsigned id = 0; id < 256; ++id)
	{
		const MaterialMap::Material* m = materials ? materials->GetMaterial((unsigned char)id) : nullptr;
		valid[id] = m ? 1 : 0;
		for (int k = 0; k < 3; ++k) { table[id * 6 + k] = m ? m->DiffuseIds0[k] : 0; table[id * 6 + 3 + k] = m ? m->DiffuseIds1[k] : 0; }
	}
	return vxb_set_materials(ctx, table, valid) == VXB_OK;
}
}

namespace
{
// host-side stage times of the last full Execute (milliseconds): materials, block offsets, blob packing, upload + decode,
// kernels, download + block views - read by bench.py through voxels_b200_last_execute_stages
double g_LastStages[6] = { 0, 0, 0, 0, 0, 0 };
struct StageClock
{
	std::chrono::steady_clock::time_point last = std::chrono::steady_clock::now();
	void mark(int stage)
	{
		const auto now = std::chrono::steady_clock::now();
		g_LastStages[stage] = std::chrono::duration<double, std::milli>(now - last).count();
		last = now;
	}
};
}

extern "C" void voxels_b200_last_execute_stages(double* out6) { for (int i = 0; i < 6; ++i) out6[i] = g_LastStages[i]; }

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
	// Execute(grid, materials, modification) with a surface to update: incremental path (:362-364, :429-465)
	if (modification && modification->Map)
	{
		SurfaceImpl* surface = static_cast<SurfaceImpl*>(modification->Map);
		if (surface->Context && unsigned(surface->Extents.x) == n)
		{
			if (!uploadMaterials(surface->Context, materials)) { logMessage(LS_Error, "Unable to polygonize grid: material table upload failed"); return nullptr; }
			return ExecuteIncremental(grid, surface, static_cast<ModificationImpl*>(modification));
		}
	}

	vxb_context* ctx = acquireContext();
	if (!ctx)
	{
		snprintf(buffer, VOXELS_LOG_SIZE, "Unable to polygonize grid: the B200 backend could not be initialised (%s)", vxb_last_error(nullptr));
		logMessage(LS_Error, buffer);
		return nullptr;
	}
	auto fail = [&](const char* what) -> PolygonSurface* {
		snprintf(buffer, VOXELS_LOG_SIZE, "Unable to polygonize grid: %s (%s)", what, vxb_last_error(ctx));
		logMessage(LS_Error, buffer);
		releaseContext(ctx);
		return nullptr;
	};
	StageClock clock;
	if (!uploadMaterials(ctx, materials)) return fail("material table upload failed");
	clock.mark(0);

	// ---- grid -> device: the grid store's compressed blocks, laid out as PackForSave does (VoxelGrid.cpp:269-315: header,
	// 3 sizes per block, then per block {flags, distance, material, blend}) in page-locked staging - a parallel memcpy of
	// ~0.2 bytes per voxel - copied as is and run-length decoded on the GPU (vxb_grid_upload_packed_streamed) ----
	const unsigned nb = n / 16;
	if (!m_Staging) m_Staging = acquireStaging();
	{
		const auto& blocks = grid.GetInternalRepresentation()->m_Blocks;
		const size_t count = blocks.size();
		if (count != size_t(nb) * nb * nb) return fail("the grid holds no block data (Grid::Create without a surface)");
		std::vector<size_t> offsets;
		const size_t total = blobLayout(blocks, offsets);
		if (!m_Staging->Blob.ensure(total)) return fail("pinned staging allocation failed");
		unsigned char* blob = static_cast<unsigned char*>(m_Staging->Blob.p);
		packHeader(n, blob);
		#pragma omp parallel for schedule(static) num_threads(packThreads())
		for (long b = 0; b < long(count); ++b) // the size table has to be complete before the first slab moves
		{
			const uint32_t sz[3] = { uint32_t(blocks[b].DistanceData.size()), uint32_t(blocks[b].MaterialData.size()), uint32_t(blocks[b].BlendData.size()) };
			memcpy(blob + 16 + size_t(b) * 12, sz, 12);
		}
		clock.mark(1);
		// the gather of slab k+1 (host threads) runs while slab k is copied and decoded
		struct Feed { const std::vector<VoxelGrid::Block>* blocks; const std::vector<size_t>* offsets; unsigned char* blob; size_t layer; double ms; };
		Feed feed = { &blocks, &offsets, blob, size_t(nb) * nb, 0.0 };
		vxb_pack_producer produce = [](void* user, uint32_t layer0, uint32_t layer1) {
			Feed* f = static_cast<Feed*>(user);
			const auto t0 = std::chrono::steady_clock::now();
			packBlocks(*f->blocks, *f->offsets, layer0 * f->layer, layer1 * f->layer, f->blob);
			f->ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
		};
		if (vxb_grid_upload_packed_streamed(ctx, blob, total, produce, &feed) != VXB_OK) return fail("grid upload failed");
		clock.mark(3);
		g_LastStages[2] = feed.ms;  // the gather (inside the upload call)
		g_LastStages[3] -= feed.ms; // what the upload call took on top of it: the last slab's copy + decode, maps
	}

	// ---- polygonize on the device ----
	if (vxb_polygonize(ctx, 0, 0) != VXB_OK) return fail("polygonization failed");
	vxb_result_info info;
	if (vxb_result_info_get(ctx, &info) != VXB_OK) return fail("no result");
	clock.mark(4);

	SurfaceImpl* surface = new SurfaceImpl;
	surface->Extents = float3(float(n), float(n), float(n)); // (W, H, D) :481
	surface->Levels.resize(info.levels_total);
	{
		Downloaded dl;
		if (!downloadResult(ctx, n, info, dl)) { delete surface; return fail("result download failed"); }
		commitResult(surface, dl, info);
	}
	surface->Context = ctx; // from here on the surface owns the device state
	{
		// PolygonMap::GetCacheSizeBytes (:196-220): consistency bits of every level-0 block + {id, blend} of every coarser cell
		size_t total = (size_t(nb) * nb * nb * 4096) >> 3;
		for (unsigned l = 1; l < info.levels_total; ++l) { const size_t c = nb >> l; total += c * c * c * 4096 * 2; }
		surface->CacheBytes = unsigned(total);
	}
	logUnmapped(ctx);
	clock.mark(5);
	if (modification) modification->Map = surface; // (the reference would have dereferenced the null Map, :443)
	return surface;
}

namespace
{
// VXB200_TRACE=1: cumulative host-side stage times of the incremental path, printed at exit
struct StageTrace
{
	bool on = getenv("VXB200_TRACE") != nullptr;
	double seconds[5] = { 0, 0, 0, 0, 0 };
	unsigned long runs = 0;
	std::chrono::steady_clock::time_point last;
	void start() { if (on) last = std::chrono::steady_clock::now(); }
	void mark(int stage)
	{
		if (!on) return;
		const auto now = std::chrono::steady_clock::now();
		seconds[stage] += std::chrono::duration<double>(now - last).count();
		last = now;
	}
	~StageTrace()
	{
		if (on && runs)
			fprintf(stderr, "[vxb200] incremental runs %lu: read blocks %.3f ms, update %.3f ms, polygonize %.3f ms, splice %.3f ms, download %.3f ms (per run)\n",
				runs, 1e3 * seconds[0] / runs, 1e3 * seconds[1] / runs, 1e3 * seconds[2] / runs, 1e3 * seconds[3] / runs, 1e3 * seconds[4] / runs);
	}
} g_Trace;
}

PolygonSurface* TransVoxelImpl::ExecuteIncremental(const Grid& grid, SurfaceImpl* surface, ModificationImpl* modification)
{
	char buffer[VOXELS_LOG_SIZE];
	g_Trace.start();
	vxb_context* ctx = surface->Context;
	const unsigned n = grid.GetWidth(), nb = n / 16;
	auto fail = [&](const char* what) -> PolygonSurface* {
		snprintf(buffer, VOXELS_LOG_SIZE, "Unable to polygonize grid: %s (%s)", what, vxb_last_error(ctx));
		logMessage(LS_Error, buffer);
		return nullptr;
	};
	// ---- refresh the device copy of the edited region: the dirty box is in OUTPUT coordinates (x, grid z, grid y) ----
	const float3 lo = modification->MinCornerModified, hi = modification->MaxCornerModified;
	const float gmin[3] = { lo.x, lo.z, lo.y }, gmax[3] = { hi.x, hi.z, hi.y };
	unsigned b0[3], b1[3];
	for (int a = 0; a < 3; ++a)
	{
		const float mn = gmin[a] < 0.f ? 0.f : gmin[a], mx = gmax[a] < 0.f ? 0.f : gmax[a];
		b0[a] = std::min(unsigned(mn) / 16u, nb - 1);
		b1[a] = std::min(unsigned(mx) / 16u, nb - 1);
		if (b1[a] < b0[a]) b1[a] = b0[a];
	}
	const size_t count = size_t(b1[0] - b0[0] + 1) * (b1[1] - b0[1] + 1) * (b1[2] - b0[2] + 1);
	if (!m_Staging) m_Staging = acquireStaging();
	if (!m_Staging->Dist.ensure(count * 4096) || !m_Staging->Mat.ensure(count * 4096) || !m_Staging->Blend.ensure(count * 4096)) return fail("pinned staging allocation failed");
	std::vector<uint32_t> coords(count * 3);
	{
		size_t i = 0;
		for (unsigned z = b0[2]; z <= b1[2]; ++z) for (unsigned y = b0[1]; y <= b1[1]; ++y) for (unsigned x = b0[0]; x <= b1[0]; ++x, ++i)
		{
			coords[i * 3] = x; coords[i * 3 + 1] = y; coords[i * 3 + 2] = z;
			const float3 c = float3(float(x), float(y), float(z));
			grid.GetBlockDistanceData(c, static_cast<char*>(m_Staging->Dist.p) + i * 4096);
			grid.GetBlockMaterialData(c, static_cast<unsigned char*>(m_Staging->Mat.p) + i * 4096, static_cast<unsigned char*>(m_Staging->Blend.p) + i * 4096);
		}
	}
	g_Trace.mark(0);
	if (vxb_grid_update_blocks(ctx, uint32_t(count), coords.data(), static_cast<const int8_t*>(m_Staging->Dist.p), static_cast<const uint8_t*>(m_Staging->Mat.p),
		static_cast<const uint8_t*>(m_Staging->Blend.p)) != VXB_OK) return fail("grid update failed");

	g_Trace.mark(1);
	// ---- re-polygonize the dirty boxes of all levels ----
	const float minCorner[3] = { lo.x, lo.y, lo.z }, maxCorner[3] = { hi.x, hi.y, hi.z };
	if (vxb_polygonize_region(ctx, minCorner, maxCorner, 0) != VXB_OK) return fail("incremental polygonization failed");
	vxb_result_info info;
	vxb_region_info region;
	if (vxb_result_info_get(ctx, &info) != VXB_OK || vxb_region_info_get(ctx, &region) != VXB_OK) return fail("no result");

	g_Trace.mark(2);
	// the new blocks come to the host first: a failed download leaves the surface as it was
	Downloaded dl;
	if (!downloadResult(ctx, n, info, dl)) return fail("result download failed");
	g_Trace.mark(4);
	// ---- splice: drop the old blocks of each level's dirty box (:443-450), append the new ones (:1293) ----
	for (unsigned l = 0; l < region.levels && l < surface->Levels.size(); ++l)
	{
		const float* mn = region.min_dirty[l]; const float* mx = region.max_dirty[l];
		auto& blocks = surface->Levels[l];
		blocks.erase(std::remove_if(blocks.begin(), blocks.end(), [&](const BlockView& b) {
			const float3& c = b.MinimalCorner;
			const bool dirty = c.x >= mn[0] && c.y >= mn[1] && c.z >= mn[2] && c.x < mx[0] && c.y < mx[1] && c.z < mx[2];
			if (dirty) --b.Owner->Live;
			return dirty;
		}), blocks.end());
		for (uint32_t i = 0; i < region.block_count[l]; ++i) modification->ModifiedBlocks.push_back(region.id_start[l] + i); // :463
	}
	commitResult(surface, dl, info);
	logUnmapped(ctx);
	g_Trace.mark(3);
	++g_Trace.runs;
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
