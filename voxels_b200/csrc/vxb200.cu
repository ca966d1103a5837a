// libvxb200.so - host side of the C ABI declared in include/vxb200.h: context, device memory,
// kernel launches (CUDA stream + events), result directory.  No CPU fallback: every entry point
// needs a CUDA device.  Kernels live in vxb_kernels.cuh.
#include <cuda_runtime.h>
#include <cuda.h> // CUtensorMap types only; cuTensorMapEncodeTiled is resolved through the runtime (no libcuda link)

#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/vxb200.h"

// the Transvoxel tables twice: __constant__ (generic kernel, uniform-ish reads) and plain global memory
// (vxbG*: divergent per-thread lookups go through L1 instead of serialising in the constant cache)
#define VXB_TABLE_QUAL __constant__
#include "vxb_tables_data.h"
#undef VXB_TABLE_QUAL
#undef VXB_TABLE_NAME
#define VXB_TABLE_QUAL __device__
#define VXB_TABLE_NAME(x) vxbG##x
#include "vxb_tables_data.h"
#include "vxb_cell.h"
#include "vxb_kernels.cuh"
#include "vxb_emit.cuh"
#include "vxb_grid.cuh"
#include "vxb_draw.cuh"

typedef VxbDecideSmem<4096> VxbDecideSmemBig;

static_assert(sizeof(VxbVertex) == 48, "PolygonVertex layout");
static_assert(sizeof(vxb_block_record) == 128, "record layout");

namespace
{
std::string g_createError;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
	const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <typename T>
struct DevBuf
{
	T* p = nullptr;
	size_t count = 0;
	cudaError_t ensure(size_t want)
	{
		if (want <= count) return cudaSuccess;
		if (p) cudaFree(p);
		p = nullptr; count = 0;
		const cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&p), want * sizeof(T));
		if (e == cudaSuccess) count = want;
		return e;
	}
	void release() { if (p) cudaFree(p); p = nullptr; count = 0; }
};
}

struct vxb_context
{
	int device = 0;
	int smCount = 0;
	cudaStream_t stream = nullptr, stream2 = nullptr;
	cudaEvent_t evBegin = nullptr, evEnd = nullptr, evFork = nullptr, evDecide0 = nullptr, evJoin = nullptr, evDir = nullptr, evVerts = nullptr;
	std::vector<cudaEvent_t> evChunks; // packed upload: one per z-chunk
	std::vector<cudaEvent_t> kevents; // per-kernel timing (pairs)
	std::string error;
	EncodeTiledFn encodeTiled = nullptr;

	uint32_t n = 0;
	int levels = 0;
	bool haveGrid = false, ownsGrid = false;
	const int8_t* dDist = nullptr; const uint8_t* dMat = nullptr; const uint8_t* dBlend = nullptr;
	DevBuf<uint8_t> volDist, volMat, volBlend, staging;
	DevBuf<unsigned long long> packOffsets;
	DevBuf<float> fillColumns;
	DevBuf<vxb_draw_command> drawCmd, drawTCmd;
	DevBuf<vxb_draw_info> drawInfo, drawTInfo;
	DevBuf<unsigned int> drawCounts;
	vxb_draw_lists drawLists;
	VxbDev lastDev; // kernel-side view of the last run (records, arenas), for the consumer-side kernels
	DevBuf<unsigned int> packSizes, packFlags, packError;
	struct PinnedBuf // page-locked host scratch (grows, never shrinks)
	{
		void* p = nullptr; size_t bytes = 0;
		bool ensure(size_t want)
		{
			if (want <= bytes) return true;
			if (p) cudaFreeHost(p);
			p = nullptr; bytes = 0;
			if (cudaHostAlloc(&p, want, cudaHostAllocDefault) != cudaSuccess) { p = nullptr; return false; }
			bytes = want; return true;
		}
		void release() { if (p) cudaFreeHost(p); p = nullptr; bytes = 0; }
	} hostOffsets, hostRecords;
	DevBuf<unsigned int> scanFlags;
	DevBuf<unsigned char> blockInfo;
	DevBuf<unsigned int> consPages;
	DevBuf<unsigned char> validFlags; // consValid + cacheValid[l] packed
	DevBuf<unsigned short> cachePages;
	DevBuf<unsigned int> worklist, emitList, bigList, transList, ntScratch, cellBlock, vlist;
	DevBuf<VxbCellRec> cellRecs;
	DevBuf<uint2> tvlist;
	DevBuf<VxbBlockRec> blockRecs;
	DevBuf<VxbVertex> verts, tverts;
	DevBuf<unsigned int> idx, tidx;
	DevBuf<vxb_block_record> records;
	DevBuf<VxbCounters> counters;
	DevBuf<VxbMaterialLut> lut;
	uint64_t capV = 0, capI = 0, capTV = 0, capTI = 0;
	CUtensorMap tmap, tmap1, tmapDist19;
	int gridVertexBlock = 0;
	DevBuf<uint8_t> lattice1;
	uint8_t* latticePtr = nullptr;  // lattice1.p, or the cube's fourth channel
	bool haveLattice1 = false, latticeOff = false;
	int gridBlock[3] = { 0, 0, 0 }, gridBlock0Small = 0, level0Threads = 256, gridDecideBig = 0, gridTransition = 0;
	DevBuf<unsigned char> mixInfo, coarseDone;
	DevBuf<unsigned short> mixCount;
	// lattices of the coarse levels (VxbDev::coarseLattice): own buffer, or - sharded runs - an area of the buffer the peers map
	DevBuf<unsigned char> coarseLatticeBuf;
	DevBuf<CUtensorMap> coarseMaps;
	unsigned char* coarseLatticeBase = nullptr;
	size_t coarseLatticeOff[VXB_MAX_LEVELS] = { 0 }, coarseLatticeBytes = 0;
	int coarseLo = 2;
	uint64_t capC = 0;
	bool haveFullRun = false;      // device caches (consistency / material pages) describe the current grid
	uint32_t nextId = 0;           // PolygonMap::GetNextBlockId (TransVoxelImpl.cpp:149-152)
	vxb_region_info region;
	DevBuf<unsigned int> updCoords;
	bool directoryFetched = false;
	bool downloadOpen = false;            // between vxb_result_download_begin and _end
	void* downloadVerts = nullptr;
	void* downloadTransVerts = nullptr;
	uint32_t shardLaunches = 0;
	// sharded runs (vxb_shard_*)
	struct Shard
	{
		bool on = false;
		uint32_t rank = 0, world = 1, groupLayers = 0; // groups of groupLayers level-0 block layers, dealt cyclically
		int sbLevel = 0;                               // super-block level = coarseLo - 1
		DevBuf<unsigned char> sbMine;
		DevBuf<unsigned int> sbWeight;
		// level-sbLevel pages + valid flags in ONE buffer the peers can map: VMM allocation (real ranks) or cudaMalloc (virtual ranks)
		unsigned char* pagesBuf = nullptr; size_t pagesBytes = 0, validBytes = 0, latticeBytes = 0, bufBytes = 0; // pages | valid | coarse lattices
		unsigned char* peerLattice[8] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
		bool pagesVmm = false;
		CUdeviceptr pagesVa = 0; CUmemGenericAllocationHandle pagesHandle = 0;
		unsigned short* peerPages[8] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
		unsigned char* peerValid[8] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
		bool peerSet[8] = { false, false, false, false, false, false, false, false };
		std::vector<std::pair<CUdeviceptr, size_t> > peerMapped;
		std::vector<CUmemGenericAllocationHandle> peerHandles;
		void* comm = nullptr;                          // ncclComm_t
		DevBuf<unsigned long long> barrierBuf;         // the tiny all-gather that orders the page exchange
		uint32_t phaseLaunches = 0;
	} shard;
	// the launch sequence of a full run as a CUDA graph, re-instantiated when any kernel argument changes
	cudaGraphExec_t graphExec = nullptr;
	std::vector<unsigned char> graphKey;
	uint32_t graphLaunches = 0;
	bool graphDisabled = false;
	// the cube of a sharded run (vxb_cube_*): one virtual range per volume, this rank's slab local, peers' slabs imported
	struct Cube
	{
		bool active = false, lattice = false;
		uint32_t rank = 0, world = 1, n = 0;
		uint32_t groupPlanes = 0, pieces = 0;          // piece p = planes [p * groupPlanes, (p + 1) * groupPlanes), owner p % world
		CUdeviceptr va[4] = { 0, 0, 0, 0 };            // distance, material, blend, even lattice ((n/2)^3)
		size_t vaBytes[4] = { 0, 0, 0, 0 };
		std::vector<CUmemGenericAllocationHandle> local[4]; // this rank's pieces, in order of local index (piece = rank + k * world)
		std::vector<CUmemGenericAllocationHandle> imported;
		std::vector<std::pair<CUdeviceptr, size_t> > mapped; // every mapped piece
		size_t pieceBytes(int channel) const { return channel < 3 ? (size_t)groupPlanes * n * n : (size_t)(groupPlanes / 2) * (n / 2) * (n / 2); }
	} cube;
	VxbCounters lastCounters;
	size_t validBytes = 0;

	bool haveResult = false;
	vxb_result_info info;
	std::vector<vxb_block_record> sortedRecords;
	uint8_t lutValid[256];
	std::vector<uint8_t> unmapped; // material ids of vertices whose material had no mapping, in logging order
	float kindMs[10] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 };        // 8, 9: the two exchanges of a sharded run
	uint32_t kindLaunches[10] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 };
};

namespace
{
int fail(vxb_context* ctx, int code, const char* what, cudaError_t e = cudaSuccess)
{
	char buf[512];
	if (e != cudaSuccess) snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(e));
	else snprintf(buf, sizeof(buf), "%s", what);
	if (ctx) ctx->error = buf; else g_createError = buf;
	return code;
}

#define VXB_CUDA(ctx, call) do { const cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(ctx, VXB_ERR_CUDA, #call, e_); } while (0)

bool validSize(uint32_t n) { return n >= 16 && n <= 4096 && (n & (n - 1)) == 0; }

int levelsFor(uint32_t n) { int l = 1; for (uint32_t v = n >> 4; v >>= 1;) ++l; return l; }


int ensureGridStorage(vxb_context* ctx, uint32_t n)
{
	const size_t vol = (size_t)n * n * n;
	VXB_CUDA(ctx, ctx->volDist.ensure(vol));
	VXB_CUDA(ctx, ctx->volMat.ensure(vol));
	VXB_CUDA(ctx, ctx->volBlend.ensure(vol));
	ctx->dDist = reinterpret_cast<const int8_t*>(ctx->volDist.p);
	ctx->dMat = ctx->volMat.p; ctx->dBlend = ctx->volBlend.p;
	ctx->n = n; ctx->levels = levelsFor(n);
	// haveGrid becomes true only when an upload has completed (a failed upload leaves no half-filled grid behind)
	ctx->ownsGrid = true; ctx->haveGrid = false; ctx->haveResult = false; ctx->haveFullRun = false;
	return VXB_OK;
}


int encodeTileMap(vxb_context* ctx, CUtensorMap* map, const void* base, uint32_t n, uint32_t rows = 17, uint32_t pitch = VXB_TILE_PITCH)
{
	const cuuint64_t dims[3] = { n, n, n };
	const cuuint64_t strides[2] = { n, (cuuint64_t)n * n };
	const cuuint32_t box[3] = { pitch, rows, rows };
	const cuuint32_t estr[3] = { 1, 1, 1 };
	const CUresult r = ctx->encodeTiled(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(base), dims, strides, box, estr,
		CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (r != CUDA_SUCCESS)
	{
		char buf[128]; snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled failed: CUresult %d", (int)r);
		return fail(ctx, VXB_ERR_CUDA, buf);
	}
	return VXB_OK;
}

size_t blocksAtLevel(uint32_t n, int level) { const size_t nb = (n / 16) >> level; return nb * nb * nb; }

// coarse levels: from the first level with <= 512 blocks (but not below 2) everything runs in one launch
int coarseLoFor(uint32_t n, int levels)
{
	int lo = 2;
	while (lo < levels && blocksAtLevel(n, lo) > 512) ++lo;
	return lo;
}

#define VXB_LATTICE_LO 2 // levels >= 2 have their own sample lattice (VxbDev::coarseLattice)

// byte layout of the coarse levels' lattices in one buffer: level l at offsets[l], (h + 1)^2 rows of h + 16 bytes
size_t coarseLatticeLayout(uint32_t n, int levels, size_t offsets[VXB_MAX_LEVELS])
{
	size_t total = 0;
	for (int l = 0; l < VXB_MAX_LEVELS; ++l) offsets[l] = 0;
	for (int l = VXB_LATTICE_LO; l < levels; ++l)
	{
		const size_t h = n >> l;
		offsets[l] = total;
		total += ((h + 16) * (h + 1) * (h + 1) + 127) & ~(size_t)127;
	}
	return total;
}

// (re)builds the coarse levels' lattice area and tensor maps: base = where the lattices live
int buildCoarseLattices(vxb_context* ctx, unsigned char* sharedBase)
{
	const uint32_t n = ctx->n;
	ctx->coarseLo = coarseLoFor(n, ctx->levels);
	ctx->coarseLatticeBytes = coarseLatticeLayout(n, ctx->levels, ctx->coarseLatticeOff);
	if (sharedBase) ctx->coarseLatticeBase = sharedBase;
	else
	{
		VXB_CUDA(ctx, ctx->coarseLatticeBuf.ensure(ctx->coarseLatticeBytes ? ctx->coarseLatticeBytes : 128));
		ctx->coarseLatticeBase = ctx->coarseLatticeBuf.p;
	}
	VXB_CUDA(ctx, ctx->coarseMaps.ensure(VXB_MAX_LEVELS));
	CUtensorMap maps[VXB_MAX_LEVELS];
	memset(maps, 0, sizeof(maps));
	for (int l = VXB_LATTICE_LO; l < ctx->levels; ++l)
	{
		const cuuint64_t h = n >> l;
		const cuuint64_t dims[3] = { h + 1, h + 1, h + 1 };
		const cuuint64_t strides[2] = { h + 16, (h + 16) * (h + 1) };
		const cuuint32_t box[3] = { VXB_TILE_PITCH, 17, 17 };
		const cuuint32_t estr[3] = { 1, 1, 1 };
		const CUresult r = ctx->encodeTiled(&maps[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, ctx->coarseLatticeBase + ctx->coarseLatticeOff[l], dims, strides, box, estr,
			CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
		if (r != CUDA_SUCCESS)
		{
			char buf[128]; snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled (coarse level %d) failed: CUresult %d", l, (int)r);
			return fail(ctx, VXB_ERR_CUDA, buf);
		}
	}
	VXB_CUDA(ctx, cudaMemcpy(ctx->coarseMaps.p, maps, sizeof(maps), cudaMemcpyHostToDevice));
	return VXB_OK;
}

int buildTensorMap(vxb_context* ctx)
{
	int r = encodeTileMap(ctx, &ctx->tmap, ctx->dDist, ctx->n);
	if (r == VXB_OK) r = encodeTileMap(ctx, &ctx->tmapDist19, ctx->dDist, ctx->n, 19, VXB_DTILE_PITCH);
	if (r != VXB_OK) return r;
	// even-lattice copy for level 1 (written by vxb_scan_kernel each run); needs at least one 16-sample row.  In a sharded
	// run every rank writes the part of its own layers, so the copy must be shared like the volumes (the cube's fourth
	// channel); without that level 1 gathers its tiles from the distance volume.
	const bool cubeLattice = ctx->cube.active && ctx->cube.lattice;
	ctx->haveLattice1 = ctx->n >= 64 && !ctx->latticeOff && (!ctx->cube.active || cubeLattice);
	ctx->latticePtr = nullptr;
	if (ctx->haveLattice1)
	{
		const size_t h = ctx->n / 2;
		if (cubeLattice) ctx->latticePtr = reinterpret_cast<uint8_t*>(ctx->cube.va[3]);
		else { VXB_CUDA(ctx, ctx->lattice1.ensure(h * h * h)); ctx->latticePtr = ctx->lattice1.p; }
		r = encodeTileMap(ctx, &ctx->tmap1, ctx->latticePtr, (uint32_t)h);
	}
	else ctx->tmap1 = ctx->tmap;
	if (r == VXB_OK) r = buildCoarseLattices(ctx, ctx->shard.on ? ctx->shard.pagesBuf + ctx->shard.pagesBytes + ctx->shard.validBytes : nullptr);
	return r;
}

int finishUpload(vxb_context* ctx)
{
	const int r = buildTensorMap(ctx);
	if (r == VXB_OK) ctx->haveGrid = true;
	return r;
}

// dirty box of an incremental run, per level (GenerateBlockListForLevel :429-465): the run keeps the caches of the last
// full run and continues its block ids
struct Region
{
	int rangeMin[VXB_MAX_LEVELS][3], rangeMax[VXB_MAX_LEVELS][3]; // grid (x, y, z) block coordinates, [min, max)
	unsigned idStart[VXB_MAX_LEVELS];
	size_t count[VXB_MAX_LEVELS];
};

struct KernelTimer
{
	vxb_context* ctx; bool on; size_t used = 0;
	std::vector<int> kinds;
	void begin(int kind)
	{
		if (!on) return;
		while (ctx->kevents.size() < used + 2) { cudaEvent_t e; cudaEventCreate(&e); ctx->kevents.push_back(e); }
		cudaEventRecord(ctx->kevents[used], ctx->stream); kinds.push_back(kind);
	}
	void end() { if (!on) return; cudaEventRecord(ctx->kevents[used + 1], ctx->stream); used += 2; }
	void collect()
	{
		for (int k = 0; k < 10; ++k) ctx->kindMs[k] = 0.f;
		if (!on) return;
		for (size_t i = 0; i < kinds.size(); ++i)
		{
			float ms = 0.f; cudaEventElapsedTime(&ms, ctx->kevents[2 * i], ctx->kevents[2 * i + 1]);
			ctx->kindMs[kinds[i]] += ms;
		}
	}
};
}

// ---- the cube of a sharded run: virtual memory management through driver entry points (no libcuda link) ----
namespace
{
struct VmmApi
{
	decltype(&cuMemAddressReserve) addressReserve = nullptr;
	decltype(&cuMemAddressFree) addressFree = nullptr;
	decltype(&cuMemCreate) create = nullptr;
	decltype(&cuMemRelease) release = nullptr;
	decltype(&cuMemMap) map = nullptr;
	decltype(&cuMemUnmap) unmap = nullptr;
	decltype(&cuMemSetAccess) setAccess = nullptr;
	decltype(&cuMemExportToShareableHandle) exportHandle = nullptr;
	decltype(&cuMemImportFromShareableHandle) importHandle = nullptr;
	decltype(&cuMemGetAllocationGranularity) granularity = nullptr;
	bool ok = false;
};

const VmmApi& vmmApi()
{
	static VmmApi api;
	static bool tried = false;
	if (tried) return api;
	tried = true;
	struct Entry { const char* name; void** slot; };
	const Entry entries[] = {
		{ "cuMemAddressReserve", (void**)&api.addressReserve }, { "cuMemAddressFree", (void**)&api.addressFree },
		{ "cuMemCreate", (void**)&api.create }, { "cuMemRelease", (void**)&api.release },
		{ "cuMemMap", (void**)&api.map }, { "cuMemUnmap", (void**)&api.unmap }, { "cuMemSetAccess", (void**)&api.setAccess },
		{ "cuMemExportToShareableHandle", (void**)&api.exportHandle }, { "cuMemImportFromShareableHandle", (void**)&api.importHandle },
		{ "cuMemGetAllocationGranularity", (void**)&api.granularity },
	};
	bool all = true;
	for (const Entry& e : entries)
	{
		cudaDriverEntryPointQueryResult q;
		if (cudaGetDriverEntryPoint(e.name, e.slot, cudaEnableDefault, &q) != cudaSuccess || !*e.slot) all = false;
	}
	api.ok = all;
	return api;
}

int failCu(vxb_context* ctx, const char* what, CUresult r)
{
	char buf[160]; snprintf(buf, sizeof(buf), "%s failed: CUresult %d", what, (int)r);
	return fail(ctx, VXB_ERR_CUDA, buf);
}
#define VXB_CU(ctx, call) do { const CUresult r_ = (call); if (r_ != CUDA_SUCCESS) return failCu(ctx, #call, r_); } while (0)

void releaseCube(vxb_context* ctx)
{
	vxb_context::Cube& c = ctx->cube;
	if (!c.active) return;
	const VmmApi& api = vmmApi();
	for (const auto& m : c.mapped) api.unmap(m.first, m.second);
	for (CUmemGenericAllocationHandle h : c.imported) api.release(h);
	for (int k = 0; k < 4; ++k)
	{
		for (CUmemGenericAllocationHandle h : c.local[k]) api.release(h);
		if (c.va[k]) api.addressFree(c.va[k], c.vaBytes[k]);
	}
	c = vxb_context::Cube();
}

void releaseShard(vxb_context* ctx)
{
	vxb_context::Shard& sh = ctx->shard;
	const VmmApi& api = vmmApi();
	for (const auto& m : sh.peerMapped) { api.unmap(m.first, m.second); api.addressFree(m.first, m.second); }
	for (CUmemGenericAllocationHandle h : sh.peerHandles) api.release(h);
	if (sh.pagesVmm)
	{
		if (sh.pagesVa) { api.unmap(sh.pagesVa, sh.bufBytes); api.addressFree(sh.pagesVa, sh.bufBytes); }
		if (sh.pagesHandle) api.release(sh.pagesHandle);
	}
	else if (sh.pagesBuf) cudaFree(sh.pagesBuf);
	sh.sbMine.release(); sh.sbWeight.release(); sh.barrierBuf.release();
	void* comm = sh.comm;
	sh = vxb_context::Shard();
	sh.comm = comm; // the communicator outlives a re-configuration (vxb_destroy ends it)
}

CUmemAllocationProp slabProp(int device)
{
	CUmemAllocationProp prop;
	memset(&prop, 0, sizeof(prop));
	prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
	prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
	prop.location.id = device;
	prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
	return prop;
}

int mapRange(vxb_context* ctx, CUdeviceptr at, size_t bytes, CUmemGenericAllocationHandle handle, std::vector<std::pair<CUdeviceptr, size_t> >& mapped)
{
	const VmmApi& api = vmmApi();
	VXB_CU(ctx, api.map(at, bytes, 0, handle, 0));
	mapped.push_back(std::make_pair(at, bytes));
	CUmemAccessDesc access;
	memset(&access, 0, sizeof(access));
	access.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
	access.location.id = ctx->device;
	access.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
	VXB_CU(ctx, api.setAccess(at, bytes, &access, 1));
	return VXB_OK;
}

// ---- NCCL through dlopen (the library torch.distributed already loaded, or the system one): the data path of a
// sharded run calls ncclAllGather itself so that a whole step is ONE stream-ordered sequence (and one CUDA graph) ----
struct NcclApi
{
	int (*getUniqueId)(void*) = nullptr;
	int (*commInitRank)(void**, int, vxb_nccl_id, int) = nullptr;
	int (*allGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
	int (*commDestroy)(void*) = nullptr;
	const char* (*getErrorString)(int) = nullptr;
	bool ok = false;
};

const NcclApi& ncclApi()
{
	static NcclApi api;
	static bool tried = false;
	if (tried) return api;
	tried = true;
	void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
	if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
	if (!lib) return api;
	api.getUniqueId = (int (*)(void*))dlsym(lib, "ncclGetUniqueId");
	api.commInitRank = (int (*)(void**, int, vxb_nccl_id, int))dlsym(lib, "ncclCommInitRank");
	api.allGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(lib, "ncclAllGather");
	api.commDestroy = (int (*)(void*))dlsym(lib, "ncclCommDestroy");
	api.getErrorString = (const char* (*)(int))dlsym(lib, "ncclGetErrorString");
	api.ok = api.getUniqueId && api.commInitRank && api.allGather && api.commDestroy;
	return api;
}

int failNccl(vxb_context* ctx, const char* what, int r)
{
	char buf[200];
	const NcclApi& api = ncclApi();
	snprintf(buf, sizeof(buf), "%s failed: ncclResult %d (%s)", what, r, api.getErrorString ? api.getErrorString(r) : "?");
	return fail(ctx, VXB_ERR_CUDA, buf);
}
}

extern "C"
{

// The cube of a sharded run: pieces of `group_planes` planes, piece p owned by rank p % world.
int vxb_cube_create(vxb_context* ctx, uint32_t n, uint32_t rank, uint32_t world, uint32_t groupPlanes)
{
	if (!ctx) return VXB_ERR_ARGUMENT;
	if (!validSize(n)) return fail(ctx, VXB_ERR_ARGUMENT, "vxb_cube_create: n must be a power of two in [16, 4096]");
	if (world == 0 || world > 8 || rank >= world) return fail(ctx, VXB_ERR_ARGUMENT, "vxb_cube_create: world in [1, 8], rank < world");
	if (groupPlanes == 0) groupPlanes = n / world;
	if (groupPlanes % 32 != 0 || n % groupPlanes != 0 || (n / groupPlanes) % world != 0)
		return fail(ctx, VXB_ERR_ARGUMENT, "vxb_cube_create: group_planes must be a multiple of 32 planes that divides n into a multiple of `world` pieces");
	const VmmApi& api = vmmApi();
	if (!api.ok) return fail(ctx, VXB_ERR_CUDA, "vxb_cube_create: the driver does not export the virtual memory management entry points");
	cudaSetDevice(ctx->device);
	cudaFree(nullptr); // make sure the primary context exists and is current for the driver calls
	releaseCube(ctx);
	vxb_context::Cube& c = ctx->cube;
	c.rank = rank; c.world = world; c.n = n;
	c.groupPlanes = groupPlanes; c.pieces = n / groupPlanes;
	const CUmemAllocationProp prop = slabProp(ctx->device);
	size_t gran = 0;
	VXB_CU(ctx, api.granularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM));
	if (!gran || c.pieceBytes(0) % gran != 0)
	{
		char buf[200]; snprintf(buf, sizeof(buf), "vxb_cube_create: a piece (%zu bytes) is not a multiple of the allocation granularity (%zu bytes)", c.pieceBytes(0), gran);
		c = vxb_context::Cube();
		return fail(ctx, VXB_ERR_ARGUMENT, buf);
	}
	c.lattice = n >= 64 && c.pieceBytes(3) % gran == 0; // else level-1 tiles are gathered from the distance volume
	c.active = true;
	const int channels = c.lattice ? 4 : 3;
	for (int k = 0; k < channels; ++k)
	{
		c.vaBytes[k] = c.pieceBytes(k) * c.pieces;
		VXB_CU(ctx, api.addressReserve(&c.va[k], c.vaBytes[k], 0, 0, 0));
		for (uint32_t p = rank; p < c.pieces; p += world)
		{
			CUmemGenericAllocationHandle h = 0;
			VXB_CU(ctx, api.create(&h, c.pieceBytes(k), &prop, 0));
			c.local[k].push_back(h);
			const int r = mapRange(ctx, c.va[k] + (CUdeviceptr)(c.pieceBytes(k) * p), c.pieceBytes(k), h, c.mapped);
			if (r != VXB_OK) return r;
		}
	}
	ctx->dDist = reinterpret_cast<const int8_t*>(c.va[0]);
	ctx->dMat = reinterpret_cast<const uint8_t*>(c.va[1]);
	ctx->dBlend = reinterpret_cast<const uint8_t*>(c.va[2]);
	ctx->n = n; ctx->levels = levelsFor(n);
	ctx->ownsGrid = false; ctx->haveGrid = true; ctx->haveResult = false; ctx->haveFullRun = false;
	return buildTensorMap(ctx);
}

int vxb_cube_info(vxb_context* ctx, uint32_t* pieces, uint32_t* channels, uint64_t* pieceBytes)
{
	if (!ctx) return VXB_ERR_ARGUMENT;
	const vxb_context::Cube& c = ctx->cube;
	if (!c.active) return fail(ctx, VXB_ERR_STATE, "vxb_cube_info: no cube");
	if (pieces) *pieces = c.pieces;
	if (channels) *channels = c.lattice ? 4u : 3u;
	if (pieceBytes) for (int k = 0; k < 4; ++k) pieceBytes[k] = (k < 3 || c.lattice) ? c.pieceBytes(k) : 0;
	return VXB_OK;
}

int vxb_cube_export(vxb_context* ctx, uint32_t channel, uint32_t piece, int* fd)
{
	if (!ctx || !fd) return VXB_ERR_ARGUMENT;
	const vxb_context::Cube& c = ctx->cube;
	if (!c.active || channel >= (c.lattice ? 4u : 3u) || piece >= c.pieces || piece % c.world != c.rank)
		return fail(ctx, VXB_ERR_ARGUMENT, "vxb_cube_export: no cube, bad channel, or a piece this rank does not own");
	cudaSetDevice(ctx->device);
	int out = -1;
	VXB_CU(ctx, vmmApi().exportHandle(&out, c.local[channel][piece / c.world], CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
	*fd = out;
	return VXB_OK;
}

int vxb_cube_import(vxb_context* ctx, uint32_t channel, uint32_t piece, int fd)
{
	if (!ctx) return VXB_ERR_ARGUMENT;
	vxb_context::Cube& c = ctx->cube;
	if (!c.active || channel >= (c.lattice ? 4u : 3u) || piece >= c.pieces || piece % c.world == c.rank || fd < 0)
		return fail(ctx, VXB_ERR_ARGUMENT, "vxb_cube_import: no cube, or bad channel / piece / descriptor");
	cudaSetDevice(ctx->device);
	CUmemGenericAllocationHandle h = 0;
	VXB_CU(ctx, vmmApi().importHandle(&h, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
	c.imported.push_back(h);
	return mapRange(ctx, c.va[channel] + (CUdeviceptr)(c.pieceBytes(channel) * piece), c.pieceBytes(channel), h, c.mapped);
}

int vxb_cube_piece(vxb_context* ctx, uint32_t piece, int8_t** dist, uint8_t** mat, uint8_t** blend, uint64_t* bytes)
{
	if (!ctx) return VXB_ERR_ARGUMENT;
	const vxb_context::Cube& c = ctx->cube;
	if (!c.active || piece >= c.pieces) return fail(ctx, VXB_ERR_STATE, "vxb_cube_piece: no cube, or bad piece");
	const size_t off = c.pieceBytes(0) * piece;
	if (dist) *dist = reinterpret_cast<int8_t*>(c.va[0] + off);
	if (mat) *mat = reinterpret_cast<uint8_t*>(c.va[1] + off);
	if (blend) *blend = reinterpret_cast<uint8_t*>(c.va[2] + off);
	if (bytes) *bytes = c.pieceBytes(0);
	return VXB_OK;
}

} // extern "C"

extern "C"
{

int vxb_create(int device, vxb_context** out)
{
	if (!out) return fail(nullptr, VXB_ERR_ARGUMENT, "vxb_create: out is null");
	*out = nullptr;
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess || count == 0) return fail(nullptr, VXB_ERR_CUDA, "vxb_create: no CUDA device (this library has no CPU fallback)", e);
	if (device < 0 || device >= count) return fail(nullptr, VXB_ERR_ARGUMENT, "vxb_create: bad device index");
	e = cudaSetDevice(device);
	if (e != cudaSuccess) return fail(nullptr, VXB_ERR_CUDA, "cudaSetDevice", e);
	cudaDeviceProp prop;
	e = cudaGetDeviceProperties(&prop, device);
	if (e != cudaSuccess) return fail(nullptr, VXB_ERR_CUDA, "cudaGetDeviceProperties", e);
	if (prop.major < 10) return fail(nullptr, VXB_ERR_CUDA, "vxb_create: the kernels are built for sm_100a (Blackwell) only");

	vxb_context* ctx = new vxb_context;
	ctx->device = device; ctx->smCount = prop.multiProcessorCount;
	memset(&ctx->info, 0, sizeof(ctx->info));
	memset(&ctx->drawLists, 0, sizeof(ctx->drawLists));
	memset(&ctx->lastDev, 0, sizeof(ctx->lastDev));
	if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess ||
		(e = cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking)) != cudaSuccess ||
		(e = cudaEventCreate(&ctx->evBegin)) != cudaSuccess || (e = cudaEventCreate(&ctx->evEnd)) != cudaSuccess ||
		(e = cudaEventCreateWithFlags(&ctx->evFork, cudaEventDisableTiming)) != cudaSuccess ||
		(e = cudaEventCreateWithFlags(&ctx->evDecide0, cudaEventDisableTiming)) != cudaSuccess ||
		(e = cudaEventCreateWithFlags(&ctx->evJoin, cudaEventDisableTiming)) != cudaSuccess ||
		(e = cudaEventCreateWithFlags(&ctx->evDir, cudaEventDisableTiming)) != cudaSuccess ||
		(e = cudaEventCreateWithFlags(&ctx->evVerts, cudaEventDisableTiming)) != cudaSuccess ||
		false)
	{ fail(nullptr, VXB_ERR_CUDA, "stream/event creation", e); vxb_destroy(ctx); return VXB_ERR_CUDA; }

	void* fn = nullptr;
	cudaDriverEntryPointQueryResult qres;
	e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
	if (e != cudaSuccess || !fn) { fail(nullptr, VXB_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available", e); vxb_destroy(ctx); return VXB_ERR_CUDA; }
	ctx->encodeTiled = reinterpret_cast<EncodeTiledFn>(fn);

	ctx->graphDisabled = getenv("VXB200_NO_GRAPH") != nullptr; // plain launches (debugging, A/B timing)
	struct KernelSetup { const void* fn; size_t smem; int* grid; const char* name; int threads; };
	ctx->level0Threads = getenv("VXB200_L0_THREADS") ? atoi(getenv("VXB200_L0_THREADS")) : 256; // A/B switch: 256 (default: measured faster) or 128
	const KernelSetup setups[6] = {
		{ (const void*)vxb_block_kernel<0, 256>, sizeof(VxbBlockSmem<0, 256>), &ctx->gridBlock[0], "vxb_block_kernel<0, 256>", 256 },
		{ (const void*)vxb_block_kernel<0, 128>, sizeof(VxbBlockSmem<0, 128>), &ctx->gridBlock0Small, "vxb_block_kernel<0, 128>", 128 },
		{ (const void*)vxb_block_kernel<1, 256>, sizeof(VxbBlockSmem<1, 256>), &ctx->gridBlock[1], "vxb_block_kernel<1>", VXB_THREADS },
		{ (const void*)vxb_block_kernel<2, 256>, sizeof(VxbBlockSmem<2, 256>), &ctx->gridBlock[2], "vxb_block_kernel<2>", VXB_THREADS },
		{ (const void*)vxb_decide_kernel<4096, 1>, sizeof(VxbDecideSmemBig), &ctx->gridDecideBig, "vxb_decide_kernel<4096>", VXB_THREADS },
		{ (const void*)vxb_transition_kernel, 0, &ctx->gridTransition, "vxb_transition_kernel", VXB_TR_THREADS },
	};
	for (const KernelSetup& k : setups)
	{
		int occ = 0;
		e = cudaFuncSetAttribute(k.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k.smem);
		if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k.fn, k.threads, k.smem);
		if (e != cudaSuccess || occ < 1) { fail(nullptr, VXB_ERR_CUDA, k.name, e); vxb_destroy(ctx); return VXB_ERR_CUDA; }
		*k.grid = occ * ctx->smCount;
	}

	if (ctx->counters.ensure(1) != cudaSuccess || ctx->lut.ensure(1) != cudaSuccess) { fail(nullptr, VXB_ERR_CUDA, "cudaMalloc"); vxb_destroy(ctx); return VXB_ERR_CUDA; }
	*out = ctx;
	const int r = vxb_set_materials(ctx, nullptr, nullptr);
	if (r != VXB_OK) { g_createError = ctx->error; vxb_destroy(ctx); *out = nullptr; return r; }
	return VXB_OK;
}

void vxb_destroy(vxb_context* ctx)
{
	if (!ctx) return;
	cudaSetDevice(ctx->device);
	if (ctx->stream) cudaStreamSynchronize(ctx->stream);
	if (ctx->stream2) cudaStreamSynchronize(ctx->stream2);
	if (ctx->graphExec) cudaGraphExecDestroy(ctx->graphExec);
	ctx->hostOffsets.release(); ctx->hostRecords.release();
	releaseShard(ctx);
	if (ctx->shard.comm && ncclApi().ok) { ncclApi().commDestroy(ctx->shard.comm); ctx->shard.comm = nullptr; }
	releaseCube(ctx);
	ctx->fillColumns.release(); ctx->packSizes.release(); ctx->packFlags.release(); ctx->packError.release();
	ctx->drawCmd.release(); ctx->drawTCmd.release(); ctx->drawInfo.release(); ctx->drawTInfo.release(); ctx->drawCounts.release();
	ctx->mixInfo.release(); ctx->coarseDone.release(); ctx->mixCount.release(); ctx->coarseLatticeBuf.release(); ctx->coarseMaps.release();
	ctx->volDist.release(); ctx->volMat.release(); ctx->volBlend.release(); ctx->staging.release(); ctx->packOffsets.release(); ctx->updCoords.release(); ctx->lattice1.release();
	ctx->scanFlags.release(); ctx->blockInfo.release(); ctx->consPages.release(); ctx->validFlags.release();
	ctx->cachePages.release(); ctx->worklist.release(); ctx->emitList.release(); ctx->bigList.release(); ctx->transList.release(); ctx->ntScratch.release(); ctx->cellBlock.release(); ctx->vlist.release(); ctx->cellRecs.release(); ctx->blockRecs.release(); ctx->tvlist.release(); ctx->verts.release(); ctx->tverts.release();
	ctx->idx.release(); ctx->tidx.release(); ctx->records.release(); ctx->counters.release(); ctx->lut.release();
	for (cudaEvent_t e : ctx->kevents) cudaEventDestroy(e);
	if (ctx->evBegin) cudaEventDestroy(ctx->evBegin);
	if (ctx->evEnd) cudaEventDestroy(ctx->evEnd);
	if (ctx->evFork) cudaEventDestroy(ctx->evFork);
	if (ctx->evDecide0) cudaEventDestroy(ctx->evDecide0);
	if (ctx->evJoin) cudaEventDestroy(ctx->evJoin);
	if (ctx->evDir) cudaEventDestroy(ctx->evDir);
	if (ctx->evVerts) cudaEventDestroy(ctx->evVerts);
	for (cudaEvent_t e : ctx->evChunks) if (e) cudaEventDestroy(e);
	if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
	if (ctx->stream) cudaStreamDestroy(ctx->stream);
	delete ctx;
}

const char* vxb_last_error(const vxb_context* ctx) { return ctx ? ctx->error.c_str() : g_createError.c_str(); }

void* vxb_stream(vxb_context* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

void* vxb_exchange_stream(vxb_context* ctx) { return ctx ? (void*)ctx->stream2 : nullptr; }

int vxb_grid_upload_dense(vxb_context* ctx, uint32_t n, const int8_t* dist, const uint8_t* mat, const uint8_t* blend)
{
	if (!ctx) return VXB_ERR_ARGUMENT;
	if (!validSize(n) || !dist) return fail(ctx, VXB_ERR_ARGUMENT, "vxb_grid_upload_dense: n must be a power of two in [16, 4096] and dist non-null");
	cudaSetDevice(ctx->device);
	int r = ensureGridStorage(ctx, n);
	if (r != VXB_OK) return r;
	const size_t vol = (size_t)n * n * n;
	VXB_CUDA(ctx, cudaMemcpyAsync(ctx->volDist.p, dist, vol, cudaMemcpyHostToDevice, ctx->stream));
	if (mat) VXB_CUDA(ctx, cudaMemcpyAsync(ctx->volMat.p, mat, vol, cudaMemcpyHostToDevice, ctx->stream));
	else VXB_CUDA(ctx, cudaMemsetAsync(ctx->volMat.p, 0, vol, ctx->stream));
	if (blend) VXB_CUDA(ctx, cudaMemcpyAsync(ctx->volBlend.p, blend, vol, cudaMemcpyHostToDevice, ctx->stream));
	else VXB_CUDA(ctx, cudaMemsetAsync(ctx->volBlend.p, 0, vol, ctx->stream));
	VXB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return finishUpload(ctx);
}

int vxb_grid_upload_blocks(vxb_context* ctx, uint32_t n, const int8_t* distBlocks, const uint8_t* matBlocks, const uint8_t* blendBlocks)
{
	if (!ctx) return VXB_ERR_ARGUMENT;
	if (!validSize(n) || !distBlocks) return fail(ctx, VXB_ERR_ARGUMENT, "vxb_grid_upload_blocks: bad size or null dist");
	cudaSetDevice(ctx->device);
	int r = ensureGridStorage(ctx, n);
	if (r != VXB_OK) return r;
	const size_t vol = (size_t)n * n * n;
	VXB_CUDA(ctx, ctx->staging.ensure(vol));
	const unsigned blocks = (unsigned)(vol / 4096);
	const void* src[3] = { distBlocks, matBlocks, blendBlocks };
	uint8_t* dst[3] = { ctx->volDist.p, ctx->volMat.p, ctx->volBlend.p };
	for (int c = 0; c < 3; ++c)
	{
		if (!src[c]) { VXB_CUDA(ctx, cudaMemsetAsync(dst[c], 0, vol, ctx->stream)); continue; }
		VXB_CUDA(ctx, cudaMemcpyAsync(ctx->staging.p, src[c], vol, cudaMemcpyHostToDevice, ctx->stream));
		vxb_unpack_blocks_kernel<<<blocks, VXB_THREADS, 0, ctx->stream>>>(reinterpret_cast<const uint4*>(ctx->staging.p), dst[c], (int)n);
		VXB_CUDA(ctx, cudaGetLastError());
	}
	VXB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return finishUpload(ctx);
}

int vxb_grid_upload_packed(vxb_context* ctx, const void* blob, size_t size)
{
	return vxb_grid_upload_packed_streamed(ctx, blob, size, nullptr, nullptr);
}

int vxb_grid_upload_packed_streamed(vxb_context* ctx, const void* blob, size_t size, vxb_pack_producer produce, void* user)
{
	if (!ctx) return VXB_ERR_ARGUMENT;
	if (!blob || size < 16) return fail(ctx, VXB_ERR_ARGUMENT, "vxb_grid_upload_packed: null or truncated blob");
	const unsigned char* bytes = static_cast<const unsigned char*>(blob);
	uint32_t header[4];
	memcpy(header, bytes, 16);
	if (header[0] != 1) return fail(ctx, VXB_ERR_ARGUMENT, "vxb_grid_upload_packed: unsupported file version (VoxelGrid.cpp:227-231)");
	const uint32_t n = header[1];
	if (!validSize(n) || header[2] != n || header[3] != n) return fail(ctx, VXB_ERR_ARGUMENT, "vxb_grid_upload_packed: the grid must be a cube with a power-of-two edge");
	const size_t nb = n / 16, blocks = nb * nb * nb;
	const size_t tableBytes = blocks * 12;
	if (size < 16 + tableBytes) return fail(ctx, VXB_ERR_ARGUMENT, "vxb_grid_upload_packed: truncated size table");
	cudaSetDevice(ctx->device);
	// a context with a cube (sharded runs) receives only the pieces this rank backs: the same blob on every rank, each
	// copies and decodes the byte ranges of its own block layers
	const bool pieces = ctx->cube.active;
	if (pieces && ctx->cube.n != n) return fail(ctx, VXB_ERR_ARGUMENT, "vxb_grid_upload_packed: the blob's grid size differs from the cube's");
	int r = pieces ? VXB_OK : ensureGridStorage(ctx, n);
	if (r != VXB_OK) return r;
	VXB_CUDA(ctx, ctx->staging.ensure(size + 16));
	VXB_CUDA(ctx, ctx->packOffsets.ensure(blocks));
	const bool trace = getenv("VXB200_TRACE") != nullptr;
	auto now = [] { return std::chrono::steady_clock::now(); };
	auto msSince = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(now() - t).count(); };
	const auto t0 = now();
	// Pipeline: header + size table first; while that moves, the per-block byte offsets (prefix sum of the size table) are
	// computed on the host; then the block data in z-chunks on the copy stream, each chunk decoded on the second stream
	// as soon as it has landed, so that only the last chunk's decode is exposed.
	const size_t head = 16 + tableBytes;
	if (!ctx->hostOffsets.ensure(blocks * sizeof(unsigned long long))) return fail(ctx, VXB_ERR_CUDA, "vxb_grid_upload_packed: pinned scratch allocation failed");
	unsigned long long* const hostOffsets = static_cast<unsigned long long*>(ctx->hostOffsets.p);
	VXB_CUDA(ctx, cudaMemcpyAsync(ctx->staging.p, blob, head, cudaMemcpyHostToDevice, ctx->stream));
	bool corrupt = false, truncated = false;
	unsigned long long off = head;
	{
		// per-block byte offsets = prefix sum of the size table (sequential, 12 bytes per block).  The blocks' flag words lie
		// scattered through the blob (one cache miss each on the host), so they are checked where they are read anyway:
		// on the device, by vxb_unpack_rle_kernel (packError)
		const unsigned char* table = bytes + 16;
		for (size_t b = 0; b < blocks; ++b)
		{
			uint32_t sz[3];
			memcpy(sz, table + b * 12, 12);
			if (sz[0] > 4096 || sz[1] > 4096 || sz[2] > 4096) { corrupt = true; break; }
			hostOffsets[b] = off;
			off += 4ull + sz[0] + sz[1] + sz[2];
		}
		if (!corrupt && off > size) truncated = true;
	}
	const double msTable = msSince(t0);
	if (corrupt || truncated || off > size)
	{
		cudaStreamSynchronize(ctx->stream);
		ctx->haveGrid = false;
		return fail(ctx, VXB_ERR_ARGUMENT, corrupt ? "vxb_grid_upload_packed: corrupt size table or block flags" : "vxb_grid_upload_packed: truncated block data");
	}
	VXB_CUDA(ctx, cudaMemcpyAsync(ctx->packOffsets.p, hostOffsets, blocks * sizeof(unsigned long long), cudaMemcpyHostToDevice, ctx->stream));
	VXB_CUDA(ctx, ctx->packError.ensure(1));
	VXB_CUDA(ctx, cudaMemsetAsync(ctx->packError.p, 0, sizeof(unsigned int), ctx->stream));
	std::vector<std::pair<size_t, size_t> > ranges; // block layers [first, second)
	if (pieces)
	{
		const size_t g = ctx->cube.groupPlanes / 16;
		for (size_t p = ctx->cube.rank; p < ctx->cube.pieces; p += ctx->cube.world) ranges.push_back(std::make_pair(p * g, (p + 1) * g));
	}
	else
	{
		const int chunks = nb >= 32 && produce ? 8 : (nb >= 16 ? 4 : 1); // a producer overlaps with the copies: finer slabs
		for (int c = 0; c < chunks; ++c) ranges.push_back(std::make_pair(nb * c / chunks, nb * (c + 1) / chunks));
	}
	while (ctx->evChunks.size() < ranges.size())
	{
		cudaEvent_t ev = nullptr;
		VXB_CUDA(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
		ctx->evChunks.push_back(ev);
	}
	unsigned char* const outDist = pieces ? reinterpret_cast<unsigned char*>(ctx->cube.va[0]) : ctx->volDist.p;
	unsigned char* const outMat = pieces ? reinterpret_cast<unsigned char*>(ctx->cube.va[1]) : ctx->volMat.p;
	unsigned char* const outBlend = pieces ? reinterpret_cast<unsigned char*>(ctx->cube.va[2]) : ctx->volBlend.p;
	const size_t groupsPerLayer = ((nb + 7) / 8) * nb;
	for (size_t c = 0; c < ranges.size(); ++c)
	{
		const size_t layer0 = ranges[c].first, layer1 = ranges[c].second;
		const unsigned long long byte0 = hostOffsets[layer0 * nb * nb], byte1 = (layer1 == nb) ? off : hostOffsets[layer1 * nb * nb];
		if (produce) produce(user, (uint32_t)layer0, (uint32_t)layer1); // the caller fills this slab while the previous one moves
		VXB_CUDA(ctx, cudaMemcpyAsync(ctx->staging.p + byte0, bytes + byte0, (size_t)(byte1 - byte0), cudaMemcpyHostToDevice, ctx->stream));
		VXB_CUDA(ctx, cudaEventRecord(ctx->evChunks[c], ctx->stream));
		VXB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream2, ctx->evChunks[c], 0));
		vxb_unpack_rle_kernel<<<(unsigned)(groupsPerLayer * (layer1 - layer0)), VXB_THREADS, 0, ctx->stream2>>>(ctx->staging.p, ctx->packOffsets.p,
			reinterpret_cast<const unsigned int*>(ctx->staging.p + 16), outDist, outMat, outBlend, (int)n, (int)layer0, ctx->packError.p);
	}
	VXB_CUDA(ctx, cudaGetLastError());
	VXB_CUDA(ctx, cudaEventRecord(ctx->evJoin, ctx->stream2));
	VXB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->evJoin, 0));
	unsigned int packError = 0;
	VXB_CUDA(ctx, cudaMemcpyAsync(&packError, ctx->packError.p, sizeof(packError), cudaMemcpyDeviceToHost, ctx->stream));
	VXB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	if (packError)
	{
		ctx->haveGrid = false;
		return fail(ctx, VXB_ERR_ARGUMENT, "vxb_grid_upload_packed: malformed block data (a raw channel must hold 4096 bytes, a run-length coded one whole (length, value) pairs whose lengths add up to 4096)");
	}
	const double msKernel = msSince(t0);
	if (pieces) { ctx->haveResult = false; ctx->haveFullRun = false; r = VXB_OK; }
	else r = finishUpload(ctx);
	if (trace) fprintf(stderr, "[vxb200] upload_packed: offsets ready at %.3f ms, copy + decode done at %.3f ms, maps at %.3f ms\n", msTable, msKernel, msSince(t0));
	return r;
}

size_t vxb_pack_dense_bound(uint32_t n)
{
	const size_t nb = n / 16, blocks = nb * nb * nb;
	return 16 + blocks * 12 + blocks * (4 + 3 * 4096);
}

namespace
{
// VoxelGrid::CompressBlock (VoxelGrid.cpp:610-672): returns the stored size; raw 4096 bytes when RLE would be larger
unsigned compressChannel(const unsigned char* data, unsigned char* out, bool signedValues, bool* uncompressed, bool* isEmpty)
{
	unsigned counter = 0, size = 1;
	const int initial = signedValues ? (int)(signed char)data[0] : (int)data[0];
	unsigned char last = data[0];
	bool empty = true, effective = true;
	unsigned ctrPos = 0;
	for (unsigned i = 0; i < 4096; ++i)
	{
		const unsigned char cur = data[i];
		if (last == cur && counter < 0xFF) { ++counter; continue; }
		out[ctrPos] = (unsigned char)counter;
		out[size++] = last;
		ctrPos = size++;
		counter = 1;
		last = cur;
		const int lv = signedValues ? (int)(signed char)last : (int)last;
		if (initial * lv <= 0) empty = false;
		if (size > 4096) { effective = false; break; }
	}
	if (effective)
	{
		out[ctrPos] = (unsigned char)counter;
		out[size++] = last;
		*uncompressed = false;
		if (isEmpty) *isEmpty = empty;
		return size;
	}
	memcpy(out, data, 4096);
	*uncompressed = true;
	if (isEmpty) *isEmpty = false;
	return 4096;
}
}

int vxb_pack_dense(uint32_t n, const int8_t* dist, const uint8_t* mat, const uint8_t* blend, void* out, size_t capacity, size_t* written)
{
	if (!validSize(n) || !dist || !mat || !blend || !out) return VXB_ERR_ARGUMENT;
	const size_t nb = n / 16, blocks = nb * nb * nb;
	if (capacity < 16 + blocks * 12) return VXB_ERR_CAPACITY;
	unsigned char* bytes = static_cast<unsigned char*>(out);
	const uint32_t header[4] = { 1u, n, n, n };
	memcpy(bytes, header, 16);
	// pass 1 (parallel): compress every block into a private slot of a scratch area; pass 2: concatenate
	std::vector<unsigned char> scratch(blocks * (4 + 3 * 4096));
	std::vector<uint32_t> sizes(blocks * 3);
	const uint8_t* chans[3] = { reinterpret_cast<const uint8_t*>(dist), mat, blend };
	#pragma omp parallel for schedule(static)
	for (long b = 0; b < (long)blocks; ++b)
	{
		const size_t bx = b % nb, by = (b / nb) % nb, bz = b / (nb * nb);
		unsigned char raw[4096];
		unsigned char* slot = scratch.data() + (size_t)b * (4 + 3 * 4096);
		uint32_t flags = 0;
		unsigned pos = 4;
		for (int ch = 0; ch < 3; ++ch)
		{
			for (int z = 0; z < 16; ++z) for (int y = 0; y < 16; ++y)
				memcpy(raw + z * 256 + y * 16, chans[ch] + ((bz * 16 + z) * n + by * 16 + y) * n + bx * 16, 16);
			bool uncompressed = false, isEmpty = false;
			const unsigned sz = compressChannel(raw, slot + pos, ch == 0, &uncompressed, ch == 0 ? &isEmpty : nullptr);
			if (uncompressed) flags |= 2u << ch;         // BF_DistanceUncompressed / Material / Blend (VoxelGrid.h:70-79)
			if (ch == 0 && isEmpty) flags |= 1u;         // BF_Empty
			sizes[b * 3 + ch] = sz;
			pos += sz;
		}
		memcpy(slot, &flags, 4);
	}
	memcpy(bytes + 16, sizes.data(), blocks * 12);
	size_t off = 16 + blocks * 12;
	for (size_t b = 0; b < blocks; ++b)
	{
		const size_t len = 4ull + sizes[b * 3] + sizes[b * 3 + 1] + sizes[b * 3 + 2];
		if (off + len > capacity) return VXB_ERR_CAPACITY;
		memcpy(bytes + off, scratch.data() + b * (4 + 3 * 4096), len);
		off += len;
	}
	if (written) *written = off;
	return VXB_OK;
}

int vxb_grid_fill(vxb_context* ctx, uint32_t n, const vxb_surface* surface, const float start[3], float step)
{
	if (!ctx) return VXB_ERR_ARGUMENT;
	if (!validSize(n) || !surface || surface->kind > VXB_SURFACE_TERRAIN) return fail(ctx, VXB_ERR_ARGUMENT, "vxb_grid_fill: bad size or surface");
	cudaSetDevice(ctx->device);
	const bool pieces = ctx->cube.active;
	if (pieces && ctx->cube.n != n) return fail(ctx, VXB_ERR_ARGUMENT, "vxb_grid_fill: n differs from the cube's");
	int r = pieces ? VXB_OK : ensureGridStorage(ctx, n);
	if (r != VXB_OK) return r;
	VxbFillArgs* args = new VxbFillArgs;
	memset(args, 0, sizeof(*args));
	args->surface = *surface;
	vxs_permutation(surface->seed, args->perm);
	for (int k = 0; k < 3; ++k) args->start[k] = start ? start[k] : 0.f;
	args->step = step;
	args->n = (int)n;
	signed char* dDist = const_cast<signed char*>(reinterpret_cast<const signed char*>(pieces ? ctx->dDist : reinterpret_cast<const int8_t*>(ctx->volDist.p)));
	unsigned char* dMat = pieces ? const_cast<unsigned char*>(ctx->dMat) : ctx->volMat.p;
	unsigned char* dBlend = pieces ? const_cast<unsigned char*>(ctx->dBlend) : ctx->volBlend.p;
	if (surface->kind == VXB_SURFACE_TERRAIN)
	{
		if (ctx->fillColumns.ensure((size_t)n * n * 4) != cudaSuccess) { delete args; return fail(ctx, VXB_ERR_CUDA, "vxb_grid_fill: column buffer allocation failed"); }
		args->columns = ctx->fillColumns.p;
		vxb_fill_columns_kernel<<<(unsigned)(((size_t)n * n + 255) / 256), 256, 0, ctx->stream>>>(*args);
	}
	std::vector<std::pair<int, int> > ranges;
	if (pieces) for (uint32_t p = ctx->cube.rank; p < ctx->cube.pieces; p += ctx->cube.world) ranges.push_back(std::make_pair((int)(p * ctx->cube.groupPlanes), (int)((p + 1) * ctx->cube.groupPlanes)));
	else ranges.push_back(std::make_pair(0, (int)n));
	for (const auto& rg : ranges)
	{
		args->z0 = rg.first; args->z1 = rg.second;
		vxb_fill_kernel<<<(unsigned)ctx->smCount * 16, 256, 0, ctx->stream>>>(*args, dDist, dMat, dBlend);
	}
	const cudaError_t e = cudaGetLastError();
	const cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
	delete args;
	if (e != cudaSuccess || e2 != cudaSuccess) return fail(ctx, VXB_ERR_CUDA, "vxb_grid_fill", e != cudaSuccess ? e : e2);
	if (pieces) { ctx->haveResult = false; ctx->haveFullRun = false; return VXB_OK; }
	return finishUpload(ctx);
}

namespace
{
// touched block range + the box the reference returns (VoxelGrid.cpp:479-487 / :577-584: "DX-style" Y-up)
int prepareEdit(vxb_context* ctx, const char* who, const float position[3], const float extents[3], VxbEditArgs& a, float outMin[3], float outMax[3])
{
	if (!ctx->haveGrid || !ctx->ownsGrid) return fail(ctx, VXB_ERR_STATE, who);
	const float n = (float)ctx->n;
	memset(&a, 0, sizeof(a));
	a.n = (int)ctx->n;
	const int nb = (int)(ctx->n / 16);
	bool any = true;
	for (int k = 0; k < 3; ++k)
	{
		a.position[k] = position[k]; a.extents[k] = extents[k];
		const float lo = position[k] - extents[k] / 2.f, hi = lo + extents[k];
		// blocks whose section is non-empty: the ones the half-open range [lo, hi) clipped to the grid reaches
		const float clo = lo < 0.f ? 0.f : lo, chi = hi > n ? n : hi;
		if (!(chi > clo)) any = false;
		int b0 = (int)floorf(clo / 16.f), b1 = (int)ceilf(chi / 16.f);
		if (b0 < 0) b0 = 0;
		if (b1 > nb) b1 = nb;
		a.b0[k] = b0; a.bn[k] = b1 > b0 ? b1 - b0 : 0;
		if (a.bn[k] == 0) any = false;
	}
	const float icp[3] = { position[0] - extents[0] / 2.f, position[1] - extents[1] / 2.f, position[2] - extents[2] / 2.f };
	const float gs[3] = { std::max(0.f, icp[0]), std::max(0.f, icp[2]), std::max(0.f, icp[1]) };
	if (outMin) { outMin[0] = gs[0]; outMin[1] = gs[1]; outMin[2] = gs[2]; }
	if (outMax) { outMax[0] = std::min(n, gs[0] + extents[0]); outMax[1] = std::min(n, gs[1] + extents[2]); outMax[2] = std::min(n, gs[2] + extents[1]); }
	return any ? VXB_OK : 1; // 1 = nothing to touch
}
}

int vxb_grid_inject_surface(vxb_context* ctx, const float position[3], const float extents[3], const vxb_surface* surface, int type, float outMin[3], float outMax[3])
{
	if (!ctx || !position || !extents || !surface) return VXB_ERR_ARGUMENT;
	if (type < 0 || type > 2 || surface->kind > VXB_SURFACE_TERRAIN) return fail(ctx, VXB_ERR_ARGUMENT, "vxb_grid_inject_surface: bad injection type or surface");
	cudaSetDevice(ctx->device);
	VxbEditArgs a;
	const int r = prepareEdit(ctx, "vxb_grid_inject_surface: needs a context-owned grid (upload or fill)", position, extents, a, outMin, outMax);
	if (r < 0) return r;
	if (r == 1) return VXB_OK;
	a.surface = *surface; a.type = type;
	if (surface->kind == VXB_SURFACE_TERRAIN) vxs_permutation(surface->seed, a.perm);
	vxb_inject_surface_kernel<<<(unsigned)(a.bn[0] * a.bn[1] * a.bn[2]), 256, 0, ctx->stream>>>(a, reinterpret_cast<signed char*>(ctx->volDist.p));
	VXB_CUDA(ctx, cudaGetLastError());
	ctx->haveResult = false;
	return VXB_OK;
}

int vxb_grid_inject_material(vxb_context* ctx, const float position[3], const float extents[3], uint32_t material, int addSubtractBlend, float outMin[3], float outMax[3])
{
	if (!ctx || !position || !extents) return VXB_ERR_ARGUMENT;
	if (material > 255) return fail(ctx, VXB_ERR_ARGUMENT, "vxb_grid_inject_material: material id > 255");
	cudaSetDevice(ctx->device);
	VxbEditArgs a;
	const int r = prepareEdit(ctx, "vxb_grid_inject_material: needs a context-owned grid (upload or fill)", position, extents, a, outMin, outMax);
	if (r < 0) return r;
	if (r == 1) return VXB_OK;
	a.material = (int)material; a.addBlend = addSubtractBlend ? 1 : 0;
	vxb_inject_material_kernel<<<(unsigned)(a.bn[0] * a.bn[1] * a.bn[2]), 256, 0, ctx->stream>>>(a, ctx->volMat.p, ctx->volBlend.p);
	VXB_CUDA(ctx, cudaGetLastError());
	ctx->haveResult = false;
	return VXB_OK;
}

int vxb_grid_pack(vxb_context* ctx, void* out, size_t capacity, size_t* written)
{
	if (!ctx || !out) return VXB_ERR_ARGUMENT;
	if (!ctx->haveGrid) return fail(ctx, VXB_ERR_STATE, "vxb_grid_pack: no grid");
	cudaSetDevice(ctx->device);
	const uint32_t n = ctx->n;
	const size_t nb = n / 16, blocks = nb * nb * nb;
	const size_t head = 16 + blocks * 12;
	if (capacity < head) return fail(ctx, VXB_ERR_CAPACITY, "vxb_grid_pack: capacity below the header + size table");
	VXB_CUDA(ctx, ctx->packSizes.ensure(blocks * 3));
	VXB_CUDA(ctx, ctx->packFlags.ensure(blocks));
	VXB_CUDA(ctx, ctx->packOffsets.ensure(blocks + 1));
	const unsigned char* d = reinterpret_cast<const unsigned char*>(ctx->dDist);
	vxb_pack_sizes_kernel<<<(unsigned)blocks, VXB_THREADS, 0, ctx->stream>>>(d, ctx->dMat, ctx->dBlend, (int)n, ctx->packSizes.p, ctx->packFlags.p);
	vxb_pack_offsets_kernel<<<1, 1024, 0, ctx->stream>>>(ctx->packSizes.p, ctx->packOffsets.p, blocks, (unsigned long long)head);
	unsigned long long total = 0;
	VXB_CUDA(ctx, cudaMemcpyAsync(&total, ctx->packOffsets.p + blocks, sizeof(total), cudaMemcpyDeviceToHost, ctx->stream));
	VXB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	if (total > capacity) return fail(ctx, VXB_ERR_CAPACITY, "vxb_grid_pack: capacity too small (vxb_pack_dense_bound is always enough)");
	VXB_CUDA(ctx, ctx->staging.ensure((size_t)total + 16));
	vxb_pack_write_kernel<<<(unsigned)blocks, VXB_THREADS, 0, ctx->stream>>>(d, ctx->dMat, ctx->dBlend, (int)n, ctx->packFlags.p, ctx->packOffsets.p, ctx->staging.p);
	VXB_CUDA(ctx, cudaGetLastError());
	unsigned char* bytes = static_cast<unsigned char*>(out);
	const uint32_t header[4] = { 1u, n, n, n };
	memcpy(bytes, header, 16);
	VXB_CUDA(ctx, cudaMemcpyAsync(bytes + 16, ctx->packSizes.p, blocks * 12, cudaMemcpyDeviceToHost, ctx->stream));
	VXB_CUDA(ctx, cudaMemcpyAsync(bytes + head, ctx->staging.p + head, (size_t)total - head, cudaMemcpyDeviceToHost, ctx->stream));
	VXB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	if (written) *written = (size_t)total;
	return VXB_OK;
}

int vxb_grid_download_dense(vxb_context* ctx, int8_t* dist, uint8_t* mat, uint8_t* blend)
{
	if (!ctx) return VXB_ERR_ARGUMENT;
	if (!ctx->haveGrid) return fail(ctx, VXB_ERR_STATE, "vxb_grid_download_dense: no grid");
	cudaSetDevice(ctx->device);
	const size_t vol = (size_t)ctx->n * ctx->n * ctx->n;
	if (dist) VXB_CUDA(ctx, cudaMemcpyAsync(dist, ctx->dDist, vol, cudaMemcpyDeviceToHost, ctx->stream));
	if (mat) VXB_CUDA(ctx, cudaMemcpyAsync(mat, ctx->dMat, vol, cudaMemcpyDeviceToHost, ctx->stream));
	if (blend) VXB_CUDA(ctx, cudaMemcpyAsync(blend, ctx->dBlend, vol, cudaMemcpyDeviceToHost, ctx->stream));
	VXB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return VXB_OK;
}

int vxb_grid_set_device(vxb_context* ctx, uint32_t n, const int8_t* dDist, const uint8_t* dMat, const uint8_t* dBlend)
{
	if (!ctx) return VXB_ERR_ARGUMENT;
	if (!validSize(n) || !dDist || !dMat || !dBlend) return fail(ctx, VXB_ERR_ARGUMENT, "vxb_grid_set_device: bad size or null pointer");
	if ((reinterpret_cast<uintptr_t>(dDist) & 15) != 0) return fail(ctx, VXB_ERR_ARGUMENT, "vxb_grid_set_device: dist must be 16-byte aligned");
	cudaSetDevice(ctx->device);
	ctx->dDist = dDist; ctx->dMat = dMat; ctx->dBlend = dBlend;
	ctx->n = n; ctx->levels = levelsFor(n);
	ctx->ownsGrid = false; ctx->haveGrid = true; ctx->haveResult = false; ctx->haveFullRun = false;
	return buildTensorMap(ctx);
}

int vxb_grid_device_pointers(vxb_context* ctx, const int8_t** dDist, const uint8_t** dMat, const uint8_t** dBlend)
{
	if (!ctx || !ctx->haveGrid) return fail(ctx, VXB_ERR_STATE, "no grid");
	if (dDist) *dDist = ctx->dDist;
	if (dMat) *dMat = ctx->dMat;
	if (dBlend) *dBlend = ctx->dBlend;
	return VXB_OK;
}

int vxb_set_materials(vxb_context* ctx, const uint8_t* table, const uint8_t* valid)
{
	if (!ctx) return VXB_ERR_ARGUMENT;
	cudaSetDevice(ctx->device);
	VxbMaterialLut lut;
	for (unsigned i = 0; i < 256; ++i)
	{
		unsigned d0[3], d1[3];
		for (int k = 0; k < 3; ++k) { d0[k] = table ? table[i * 6 + k] : i; d1[k] = table ? table[i * 6 + 3 + k] : i; }
		// byte order of PolygonVertex::Textures: Reserved, Blend, Uxz, Txz | Uny, Upy, Tny, Tpy  (TransVoxelImpl.cpp:1253-1261)
		lut.tex0[i] = (d1[1] << 16) | (d0[1] << 24);
		lut.tex1[i] = d1[2] | (d1[0] << 8) | (d0[2] << 16) | (d0[0] << 24);
		lut.valid[i] = valid ? (valid[i] != 0) : 1;
	}
	VXB_CUDA(ctx, cudaMemcpy(ctx->lut.p, &lut, sizeof(lut), cudaMemcpyHostToDevice));
	memcpy(ctx->lutValid, lut.valid, 256);
	return VXB_OK;
}

int vxb_set_capacity(vxb_context* ctx, uint64_t v, uint64_t i, uint64_t tv, uint64_t ti)
{
	if (!ctx) return VXB_ERR_ARGUMENT;
	if (v) ctx->capV = v;
	if (i) ctx->capI = i;
	if (tv) ctx->capTV = tv;
	if (ti) ctx->capTI = ti;
	return VXB_OK;
}

// shardPhase: -1 = not a sharded run; 0 = scan + block info (no sync); 1 = plan + the levels below the coarse ones + page
// publication (no sync); 2 = coarse levels, emission, directory (syncs, result); 3 = everything, the two exchanges by
// ncclAllGather on the context's stream (needs vxb_shard_nccl_init)
static int runPolygonize(vxb_context* ctx, uint32_t maxLevels, uint32_t flags, const Region* region, int shardPhase = -1)
{
	const bool incremental = region != nullptr;
	if (!ctx) return VXB_ERR_ARGUMENT;
	if (!ctx->haveGrid) return fail(ctx, VXB_ERR_STATE, "vxb_polygonize: no grid uploaded");
	if (ctx->downloadOpen) return fail(ctx, VXB_ERR_STATE, "vxb_polygonize: a result download is open (vxb_result_download_end first)");
	cudaSetDevice(ctx->device);
	ctx->haveResult = false;
	const uint32_t n = ctx->n;
	const int levels = ctx->levels;
	const int computed = (maxLevels == 0 || (int)maxLevels > levels) ? levels : (int)maxLevels;
	const size_t nb0 = n / 16, blocks0 = nb0 * nb0 * nb0;
	const bool sharded = shardPhase >= 0;
	vxb_context::Shard& sh = ctx->shard;
	const bool kernelTimes = (flags & VXB_FLAG_KERNEL_TIMES) != 0 && (!sharded || shardPhase == 3);

	// ---- device state ----
	VXB_CUDA(ctx, ctx->scanFlags.ensure(blocks0));
	VXB_CUDA(ctx, ctx->blockInfo.ensure(blocks0));
	VXB_CUDA(ctx, ctx->consPages.ensure(blocks0 * 128));
	size_t totalBlocks = 0, validBytes = 0, cacheEntries = 0, mixBytes = 0;
	size_t validOff[VXB_MAX_LEVELS], cacheOff[VXB_MAX_LEVELS], mixOff[VXB_MAX_LEVELS];
	VxbDev dev;
	memset(&dev, 0, sizeof(dev));
	const int coarseLo = ctx->coarseLo;
	size_t coarseBlocks = 0;
	for (int l = 0; l < levels; ++l)
	{
		const size_t b = blocksAtLevel(n, l);
		dev.workBase[l] = (unsigned)totalBlocks;
		dev.idBase[l] = (unsigned)totalBlocks;
		totalBlocks += b;
		validOff[l] = validBytes; validBytes += (b + 15) & ~(size_t)15;
		cacheOff[l] = cacheEntries; if (l >= 1) cacheEntries += b * 4096;
		mixOff[l] = mixBytes; if (l >= 1) mixBytes += (b + 15) & ~(size_t)15;
		dev.coarseBase[l] = (unsigned)coarseBlocks; if (l >= coarseLo) coarseBlocks += b;
	}
	VXB_CUDA(ctx, ctx->validFlags.ensure(validBytes));
	VXB_CUDA(ctx, ctx->cachePages.ensure(cacheEntries ? cacheEntries : 1));
	VXB_CUDA(ctx, ctx->mixInfo.ensure(mixBytes ? mixBytes : 16));
	VXB_CUDA(ctx, ctx->mixCount.ensure(mixBytes ? mixBytes : 16));
	VXB_CUDA(ctx, ctx->coarseDone.ensure(coarseBlocks ? coarseBlocks : 16));
	VXB_CUDA(ctx, ctx->worklist.ensure(totalBlocks));
	VXB_CUDA(ctx, ctx->records.ensure(totalBlocks));
	VXB_CUDA(ctx, ctx->emitList.ensure(totalBlocks));
	VXB_CUDA(ctx, ctx->bigList.ensure(totalBlocks));
	VXB_CUDA(ctx, ctx->transList.ensure(totalBlocks));
	VXB_CUDA(ctx, ctx->blockRecs.ensure(totalBlocks));
	VXB_CUDA(ctx, ctx->ntScratch.ensure(totalBlocks * 256));

	const uint64_t vol = sharded ? (uint64_t)n * n * n / sh.world * 2 : (uint64_t)n * n * n;
	if (!ctx->capV) ctx->capV = std::max<uint64_t>(1u << 20, vol / 24);
	if (!ctx->capI) ctx->capI = ctx->capV * 6;
	if (!ctx->capTV) ctx->capTV = std::max<uint64_t>(1u << 18, ctx->capV / 8);
	if (!ctx->capTI) ctx->capTI = ctx->capTV * 6;
	if (!ctx->capC) ctx->capC = ctx->capV + ctx->capV / 4;

	dev.grid.dist = ctx->dDist; dev.grid.mat = ctx->dMat; dev.grid.blend = ctx->dBlend; dev.grid.n = (int)n;
	dev.n = (int)n; dev.levels = levels; dev.lastLevel = levels - 1; dev.computed = computed;
	dev.scanFlags = ctx->scanFlags.p; dev.blockInfo = ctx->blockInfo.p;
	dev.consPages = ctx->consPages.p;
	dev.consValid = ctx->validFlags.p + validOff[0];
	for (int l = 1; l < levels; ++l)
	{
		dev.cachePages[l] = ctx->cachePages.p + cacheOff[l]; dev.cacheValid[l] = ctx->validFlags.p + validOff[l];
		dev.mixInfo[l] = ctx->mixInfo.p + mixOff[l];
		dev.mixCount[l] = ctx->mixCount.p + mixOff[l];
	}
	dev.coarseLo = coarseLo; dev.coarseDone = ctx->coarseDone.p;
	VxbCoarseLattices scanLat;
	memset(&scanLat, 0, sizeof(scanLat));
	scanLat.lo = VXB_LATTICE_LO; scanLat.levels = levels;
	for (int l = VXB_LATTICE_LO; l < levels; ++l) { dev.coarseLattice[l] = ctx->coarseLatticeBase + ctx->coarseLatticeOff[l]; scanLat.p[l] = dev.coarseLattice[l]; }
	dev.coarseMaps = ctx->coarseMaps.p;
	dev.latticeLo = VXB_LATTICE_LO;
	dev.worklist = ctx->worklist.p;
	dev.records = ctx->records.p; dev.rcap = (unsigned)totalBlocks;
	dev.counters = ctx->counters.p; dev.lut = ctx->lut.p;
	dev.transitions = (flags & VXB_FLAG_NO_TRANSITIONS) ? 0 : 1;
	dev.emitList = ctx->emitList.p; dev.bigList = ctx->bigList.p; dev.transList = ctx->transList.p; dev.ntScratch = ctx->ntScratch.p;
	dev.blockRecs = ctx->blockRecs.p;
	dev.lattice1 = ctx->haveLattice1 ? ctx->latticePtr : nullptr;
	dev.incremental = incremental ? 1 : 0;
	dev.ranged = region ? 1 : 0;
	if (region)
		for (int l = 0; l < levels; ++l)
		{
			for (int a = 0; a < 3; ++a) { dev.rangeMin[l][a] = region->rangeMin[l][a]; dev.rangeMax[l][a] = region->rangeMax[l][a]; }
			dev.idStart[l] = region->idStart[l];
		}
	VxbPeers peers;
	memset(&peers, 0, sizeof(peers));
	VxbPeerLattices peerLat;
	memset(&peerLat, 0, sizeof(peerLat));
	if (sharded)
	{
		if (!sh.on) return fail(ctx, VXB_ERR_STATE, "vxb_polygonize_sharded: call vxb_shard_configure first");
		if (sh.sbLevel != coarseLo - 1 || coarseLo >= levels) return fail(ctx, VXB_ERR_STATE, "vxb_polygonize_sharded: the grid changed since vxb_shard_configure");
		dev.shardWorld = (int)sh.world; dev.shardRank = (int)sh.rank; dev.shardLayers = (int)sh.groupLayers;
		dev.sbLevel = sh.sbLevel; dev.sbMine = sh.sbMine.p; dev.sbWeight = sh.sbWeight.p;
		// the pages of the super-block level live in the buffer the peers write into
		dev.cachePages[sh.sbLevel] = reinterpret_cast<unsigned short*>(sh.pagesBuf);
		dev.cacheValid[sh.sbLevel] = sh.pagesBuf + sh.pagesBytes;
		for (uint32_t p = 0; p < sh.world; ++p)
		{
			if (p == sh.rank) continue;
			if (!sh.peerSet[p]) return fail(ctx, VXB_ERR_STATE, "vxb_polygonize_sharded: a peer's page buffer is not set (vxb_shard_set_peer / vxb_shard_import)");
			peers.pages[peers.count] = sh.peerPages[p]; peers.valid[peers.count] = sh.peerValid[p];
			peers.peerSlots[peers.count] = reinterpret_cast<unsigned long long*>(sh.peerLattice[p] + sh.latticeBytes);
			peerLat.base[peers.count] = sh.peerLattice[p];
			++peers.count;
		}
		peerLat.count = peers.count;
		{
			unsigned char* area = sh.pagesBuf + sh.pagesBytes + sh.validBytes + sh.latticeBytes;
			peers.mySlots = reinterpret_cast<unsigned long long*>(area);
			peers.epoch = reinterpret_cast<unsigned long long*>(area + 64);
			peers.arrived = reinterpret_cast<unsigned int*>(area + 128);
			peers.rank = (int)sh.rank;
		}
		if (shardPhase == 3 && sh.world > 1 && (!sh.comm || !ncclApi().ok)) return fail(ctx, VXB_ERR_STATE, "vxb_polygonize_sharded: phase 3 needs vxb_shard_nccl_init");
	}
	const int scanWorld = sharded ? (int)sh.world : 1, scanRank = sharded ? (int)sh.rank : 0, scanGroup = sharded ? (int)sh.groupLayers : (int)nb0;

	VxbCounters hc;
	for (int attempt = 0; attempt < 6; ++attempt)
	{
		const uint64_t lim = 0xFFFFFFF0ull;
		ctx->capV = std::min(ctx->capV, lim); ctx->capI = std::min(ctx->capI, lim);
		ctx->capTV = std::min(ctx->capTV, lim); ctx->capTI = std::min(ctx->capTI, lim);
		ctx->capV = std::min<uint64_t>(ctx->capV, 0x0FFFFFF0ull); ctx->capC = std::min<uint64_t>(ctx->capC, 0x0FFFFFF0ull); // 28-bit cell index in the vertex list
		VXB_CUDA(ctx, ctx->verts.ensure(ctx->capV));
		VXB_CUDA(ctx, ctx->idx.ensure(ctx->capI));
		VXB_CUDA(ctx, ctx->tverts.ensure(ctx->capTV));
		VXB_CUDA(ctx, ctx->tidx.ensure(ctx->capTI));
		VXB_CUDA(ctx, ctx->vlist.ensure(ctx->capV));
		VXB_CUDA(ctx, ctx->tvlist.ensure(ctx->capTV));
		VXB_CUDA(ctx, ctx->cellRecs.ensure(ctx->capC));
		VXB_CUDA(ctx, ctx->cellBlock.ensure(ctx->capC));
		dev.verts = ctx->verts.p; dev.idx = ctx->idx.p; dev.tverts = ctx->tverts.p; dev.tidx = ctx->tidx.p;
		dev.vcap = (unsigned)ctx->capV; dev.icap = (unsigned)ctx->capI; dev.tvcap = (unsigned)ctx->capTV; dev.ticap = (unsigned)ctx->capTI;
		dev.vlist = ctx->vlist.p; dev.cellRecs = ctx->cellRecs.p; dev.cellBlock = ctx->cellBlock.p; dev.ccap = (unsigned)ctx->capC; dev.tvlist = ctx->tvlist.p;

		KernelTimer timer{ ctx, kernelTimes };
		uint32_t launches = 0;
		for (int k = 0; k < 10; ++k) ctx->kindLaunches[k] = 0;
		const unsigned flatGrid = (unsigned)ctx->smCount * 8;
		cudaStream_t st = ctx->stream;

		bool latticeForked = false;
		auto joinLattice = [&]() { if (latticeForked) { cudaStreamWaitEvent(st, ctx->evVerts, 0); latticeForked = false; } };
		// ---- the pieces of a run; a full run is all of them in order, a sharded run has an exchange after the first two ----
		auto scanPart = [&]() -> int
		{
			VXB_CUDA(ctx, cudaMemsetAsync(ctx->counters.p, 0, sizeof(VxbCounters), st));
			if (!incremental) VXB_CUDA(ctx, cudaMemsetAsync(ctx->validFlags.p, 0, validBytes, st)); // incremental runs keep the caches (:362-364)
			if (sharded) VXB_CUDA(ctx, cudaMemsetAsync(sh.pagesBuf + sh.pagesBytes, 0, sh.validBytes, st));
			if (coarseBlocks) VXB_CUDA(ctx, cudaMemsetAsync(ctx->coarseDone.p, 0, coarseBlocks, st));
			// incremental runs: only the block layers of the level-0 dirty box changed since the last run
			const int zBase = incremental ? std::max(0, region->rangeMin[0][2]) : 0;
			const size_t myLayers = incremental ? (size_t)std::max(0, std::min((int)nb0, region->rangeMax[0][2]) - zBase) : nb0 / scanWorld;
			if (!myLayers) return VXB_OK;
			const int perCta = nb0 >= 32 ? 32 : 8;
			const dim3 grid((unsigned)((nb0 + perCta - 1) / perCta), (unsigned)nb0, (unsigned)myLayers);
			timer.begin(0);
			if (perCta == 32) vxb_scan_kernel<32><<<grid, VXB_THREADS, 0, st>>>(ctx->dDist, (int)n, ctx->scanFlags.p, ctx->haveLattice1 ? ctx->latticePtr : nullptr, scanGroup, scanWorld, scanRank, zBase);
			else vxb_scan_kernel<8><<<grid, VXB_THREADS, 0, st>>>(ctx->dDist, (int)n, ctx->scanFlags.p, ctx->haveLattice1 ? ctx->latticePtr : nullptr, scanGroup, scanWorld, scanRank, zBase);
			timer.end(); ++launches; ++ctx->kindLaunches[0];
			if (VXB_LATTICE_LO < computed)
			{
				// the sample lattices of levels >= 2: needed only by vxb_block_kernel<2>, so a single-GPU run computes them on the
				// second stream next to the block walk and the level-0 kernel; a sharded rank does the rows of its own planes here
				const bool cubeRows = sharded && ctx->cube.active; // the volumes are split over the ranks: own rows here, then published
				cudaStream_t ls = (cubeRows || kernelTimes || shardPhase == 0) ? st : ctx->stream2;
				if (ls != st)
				{
					VXB_CUDA(ctx, cudaEventRecord(ctx->evDecide0, st));
					VXB_CUDA(ctx, cudaStreamWaitEvent(ls, ctx->evDecide0, 0));
				}
				timer.begin(1);
				vxb_coarse_lattice_kernel<<<(unsigned)ctx->smCount * 4, 256, 0, ls>>>(reinterpret_cast<const signed char*>(ctx->dDist), (int)n, scanLat,
					cubeRows ? (int)sh.groupLayers * 16 : 1, cubeRows ? (int)sh.world : 0, cubeRows ? (int)sh.rank : 0);
				timer.end(); ++launches; ++ctx->kindLaunches[1];
				if (ls != st) { VXB_CUDA(ctx, cudaEventRecord(ctx->evVerts, ls)); latticeForked = true; } // joined before the first kernel that reads a lattice
			}
			if (sharded && ctx->cube.active && peerLat.count && VXB_LATTICE_LO < levels)
			{
				// the lattice planes of my layers -> every peer (ordered before the peers' coarse levels by the exchanges)
				timer.begin(8);
				vxb_publish_lattice_kernel<<<(unsigned)ctx->smCount, VXB_THREADS, 0, st>>>(dev, peerLat, ctx->coarseLatticeBase);
				timer.end(); ++launches; ++ctx->kindLaunches[8];
			}
			const size_t mine = myLayers * nb0 * nb0;
			const unsigned g2 = (unsigned)std::min<size_t>((mine + 255) / 256, (size_t)ctx->smCount * 8);
			timer.begin(1);
			vxb_block_info_kernel<<<g2, 256, 0, st>>>(ctx->dDist, (int)n, ctx->scanFlags.p, ctx->blockInfo.p, mine, scanGroup, scanWorld, scanRank, zBase);
			timer.end(); ++launches; ++ctx->kindLaunches[1];
			return VXB_OK;
		};
		auto selectLevel = [&](int l) {
			const size_t b = region ? region->count[l] : blocksAtLevel(n, l);
			const unsigned gs = (unsigned)std::min<size_t>((b + 255) / 256, (size_t)ctx->smCount * 8);
			timer.begin(1);
			vxb_select_kernel<<<gs, 256, 0, st>>>(dev, l);
			timer.end(); ++launches; ++ctx->kindLaunches[1];
			return b;
		};
		auto levelPart = [&]() -> int
		{
			if (levels > 1)
			{
				// sign-mix pyramid: the block walk of levels >= 2 and (sharded runs) the weights of the super-blocks
				size_t threadsA = 0;
				for (int l = 1; l < levels && l <= 2; ++l) threadsA += blocksAtLevel(n, l);
				timer.begin(1);
				vxb_pyramid_kernel<<<(unsigned)((threadsA + 255) / 256), 256, 0, st>>>(dev);
				if (levels > 3) { vxb_pyramid_top_kernel<<<1, 1024, 0, st>>>(dev); ++launches; ++ctx->kindLaunches[1]; }
				timer.end(); ++launches; ++ctx->kindLaunches[1];
			}
			if (sharded) { timer.begin(1); vxb_plan_kernel<<<1, 1024, 0, st>>>(dev); timer.end(); ++launches; ++ctx->kindLaunches[1]; }
			if (selectLevel(0))
			{
				timer.begin(3);
				if (ctx->level0Threads == 256) vxb_block_kernel<0, 256><<<ctx->gridBlock[0], 256, sizeof(VxbBlockSmem<0, 256>), st>>>(ctx->tmapDist19, ctx->tmapDist19, dev, 0);
				else vxb_block_kernel<0, 128><<<ctx->gridBlock0Small, 128, sizeof(VxbBlockSmem<0, 128>), st>>>(ctx->tmapDist19, ctx->tmapDist19, dev, 0);
				timer.end(); ++launches; ++ctx->kindLaunches[3];
			}
			vxb_mark_split_kernel<<<1, 1, 0, st>>>(dev); ++launches; // level 0 is complete: the flat kernels work on [split, end)
			for (int l = 1; l < coarseLo && l < computed; ++l)
			{
				if (!selectLevel(l)) continue;
				if (l >= VXB_LATTICE_LO) joinLattice();
				timer.begin(2);
				vxb_block_kernel<1, 256><<<ctx->gridBlock[1], VXB_THREADS, sizeof(VxbBlockSmem<1, 256>), st>>>(ctx->tmap, ctx->tmap1, dev, l);
				timer.end(); ++launches; ++ctx->kindLaunches[2];
			}
			if (sharded && peers.count) { timer.begin(9); vxb_publish_kernel<<<(unsigned)ctx->smCount * 4, VXB_THREADS, 0, st>>>(dev, peers); timer.end(); ++launches; ++ctx->kindLaunches[9]; }
			VXB_CUDA(ctx, cudaGetLastError());
			return VXB_OK;
		};
		auto tailPart = [&]() -> int
		{
			if (computed > coarseLo)
			{
				joinLattice(); // the lattices are complete
				timer.begin(2);
				vxb_block_kernel<2, 256><<<ctx->gridBlock[2], VXB_THREADS, sizeof(VxbBlockSmem<2, 256>), st>>>(ctx->tmap, ctx->tmap1, dev, coarseLo);
				timer.end(); ++launches; ++ctx->kindLaunches[2];
			}
			// blocks too large for the shared-memory path of vxb_block_kernel (> 1024 non-trivial cells / > 2048 level-0 vertices)
			timer.begin(2);
			vxb_decide_kernel<4096, 1><<<ctx->gridDecideBig, VXB_THREADS, sizeof(VxbDecideSmemBig), st>>>(ctx->tmap, ctx->tmap1, dev, 2);
			timer.end(); ++launches; ++ctx->kindLaunches[2];
			// vertices + triangles of everything but the level-0 blocks done inside vxb_block_kernel<0>, next to the transition
			// cells (which need the block records and the material pages only)
			const bool fork = dev.transitions && computed > 1 && !kernelTimes;
			if (fork)
			{
				VXB_CUDA(ctx, cudaEventRecord(ctx->evFork, st));
				VXB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream2, ctx->evFork, 0));
			}
			timer.begin(4);
			vxb_vertex_kernel<<<flatGrid, VXB_THREADS, 0, st>>>(dev, 1);
			timer.end(); ++launches; ++ctx->kindLaunches[4];
			timer.begin(5);
			vxb_triangle_kernel<<<flatGrid, VXB_THREADS, 0, st>>>(dev, 1);
			timer.end(); ++launches; ++ctx->kindLaunches[5];
			if (dev.transitions && computed > 1)
			{
				cudaStream_t ts = fork ? ctx->stream2 : st;
				timer.begin(6);
				vxb_transition_kernel<<<ctx->gridTransition, VXB_TR_THREADS, 0, ts>>>(dev);
				vxb_transition_vertex_kernel<<<flatGrid, VXB_THREADS, 0, ts>>>(dev);
				timer.end(); launches += 2; ctx->kindLaunches[6] += 2;
				if (fork)
				{
					VXB_CUDA(ctx, cudaEventRecord(ctx->evJoin, ctx->stream2));
					VXB_CUDA(ctx, cudaStreamWaitEvent(st, ctx->evJoin, 0));
				}
			}
			joinLattice();
			timer.begin(7);
			vxb_finish_kernel<<<(unsigned)ctx->smCount * 2, VXB_THREADS, 0, st>>>(dev);
			timer.end(); ++launches; ++ctx->kindLaunches[7];
			VXB_CUDA(ctx, cudaGetLastError());
			return VXB_OK;
		};
		auto exchange = [&](int which) -> int
		{
			if (sh.world == 1) return VXB_OK; // one rank: nothing to exchange
			const NcclApi& nccl = ncclApi();
			timer.begin(8 + which);
			++ctx->kindLaunches[8 + which];
			if (which == 0)
			{
				// all-gather of the block info: every rank's layers are one contiguous chunk (rank-major layout)
				const size_t chunk = blocks0 / sh.world;
				const int r = nccl.allGather(ctx->blockInfo.p + chunk * sh.rank, ctx->blockInfo.p, chunk, /*ncclUint8*/ 1, sh.comm, st);
				if (r != 0) return failNccl(ctx, "ncclAllGather (block info)", r);
			}
			else
			{
				// the pages were written straight into the peers' buffers by vxb_publish_kernel, whose last CTA announced this
				// step's epoch to every peer; waiting for the peers' epochs orders every rank's coarse levels after every rank's
				// publication (a device-side barrier over the mapped buffers: ~5 us instead of a ~40 us one-word all-gather)
				vxb_peer_wait_kernel<<<1, 32, 0, st>>>(peers, (int)sh.world);
				++launches;
			}
			timer.end();
			return VXB_OK;
		};
		auto enqueueAll = [&]() -> int
		{
			int r = scanPart();
			if (r == VXB_OK && shardPhase == 3) r = exchange(0);
			if (r == VXB_OK) r = levelPart();
			if (r == VXB_OK && shardPhase == 3) r = exchange(1);
			if (r == VXB_OK) r = tailPart();
			return r;
		};

		if (shardPhase == 0)
		{
			VXB_CUDA(ctx, cudaEventRecord(ctx->evBegin, st));
			const int r = scanPart();
			VXB_CUDA(ctx, cudaGetLastError());
			sh.phaseLaunches = launches;
			return r;
		}
		if (shardPhase == 1)
		{
			const int r = levelPart();
			sh.phaseLaunches += launches;
			return r;
		}
		if (shardPhase != 2) VXB_CUDA(ctx, cudaEventRecord(ctx->evBegin, st));
		// A run has no host decision inside the sequence, so it is captured once into a CUDA graph and replayed; the key is
		// every kernel argument (a re-upload into other buffers, grown arenas, other flags => new capture).
		bool replayed = false;
		if ((shardPhase == -1 || shardPhase == 3) && !incremental && !kernelTimes && !ctx->graphDisabled)
		{
			std::vector<unsigned char> key(sizeof(VxbDev) + sizeof(VxbPeers) + sizeof(VxbPeerLattices) + 3 * sizeof(CUtensorMap) + 64, 0);
			unsigned char* k = key.data();
			memcpy(k, &dev, sizeof(VxbDev)); k += sizeof(VxbDev);
			memcpy(k, &peers, sizeof(VxbPeers)); k += sizeof(VxbPeers);
			memcpy(k, &peerLat, sizeof(VxbPeerLattices)); k += sizeof(VxbPeerLattices);
			const CUtensorMap* maps[3] = { &ctx->tmap, &ctx->tmap1, &ctx->tmapDist19 };
			for (const CUtensorMap* mp : maps) { memcpy(k, mp, sizeof(CUtensorMap)); k += sizeof(CUtensorMap); }
			const void* ptrs[4] = { ctx->dDist, ctx->scanFlags.p, ctx->blockInfo.p, ctx->haveLattice1 ? ctx->latticePtr : nullptr };
			memcpy(k, ptrs, sizeof(ptrs)); k += sizeof(ptrs);
			const int scalars[6] = { computed, (int)validBytes, (int)nb0, (int)n, shardPhase, (int)flags };
			memcpy(k, scalars, sizeof(scalars));
			if (!ctx->graphExec || key != ctx->graphKey)
			{
				if (ctx->graphExec) { cudaGraphExecDestroy(ctx->graphExec); ctx->graphExec = nullptr; }
				cudaGraph_t graph = nullptr;
				cudaError_t ce = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
				if (ce == cudaSuccess)
				{
					const int rc = enqueueAll();
					ce = cudaStreamEndCapture(st, &graph);
					if (rc != VXB_OK && ce == cudaSuccess) ce = cudaErrorUnknown;
				}
				if (ce == cudaSuccess) ce = cudaGraphInstantiate(&ctx->graphExec, graph, 0);
				if (graph) cudaGraphDestroy(graph);
				if (ce != cudaSuccess) { ctx->graphExec = nullptr; ctx->graphDisabled = true; cudaGetLastError(); } // fall back to plain launches for good
				else { ctx->graphKey = key; ctx->graphLaunches = launches; }
				launches = 0;
			}
			if (ctx->graphExec)
			{
				VXB_CUDA(ctx, cudaGraphLaunch(ctx->graphExec, st));
				launches = ctx->graphLaunches;
				replayed = true;
			}
		}
		if (!replayed)
		{
			const int rc = (shardPhase == 2) ? tailPart() : enqueueAll();
			if (rc != VXB_OK) return rc;
			if (shardPhase == 2) launches += sh.phaseLaunches;
		}
		VXB_CUDA(ctx, cudaGetLastError());
		VXB_CUDA(ctx, cudaEventRecord(ctx->evEnd, st));
		VXB_CUDA(ctx, cudaMemcpyAsync(&hc, ctx->counters.p, sizeof(hc), cudaMemcpyDeviceToHost, st));
		VXB_CUDA(ctx, cudaStreamSynchronize(st));
		float ms = 0.f;
		VXB_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->evBegin, ctx->evEnd));
		timer.collect();
		ctx->info.device_ms = ms;
		ctx->info.kernel_launches = launches;

		const bool overflow = hc.vertices > ctx->capV || hc.indices > ctx->capI || hc.transVertices > ctx->capTV || hc.transIndices > ctx->capTI || hc.cells > ctx->capC;
		if (!overflow) break;
		if (attempt == 5) return fail(ctx, VXB_ERR_CAPACITY, "vxb_polygonize: output arenas overflowed after growing");
		if (hc.vertices > ctx->capV) ctx->capV = (uint64_t)hc.vertices + hc.vertices / 8 + 1024;
		if (hc.indices > ctx->capI) ctx->capI = (uint64_t)hc.indices + hc.indices / 8 + 1024;
		if (hc.transVertices > ctx->capTV) ctx->capTV = (uint64_t)hc.transVertices + hc.transVertices / 8 + 1024;
		if (hc.transIndices > ctx->capTI) ctx->capTI = (uint64_t)hc.transIndices + hc.transIndices / 8 + 1024;
		if (hc.cells > ctx->capC) ctx->capC = (uint64_t)hc.cells + hc.cells / 8 + 1024;
		// a sharded run is collective: the caller repeats it on every rank
		if (sharded) return fail(ctx, VXB_ERR_CAPACITY, "vxb_polygonize_sharded: output arenas overflowed; capacities were grown, repeat the run on every rank");
	}

	// the directory stays on the device (it is part of the result resident in HBM); vxb_result_download fetches and
	// sorts it into the reference's block order on demand
	ctx->lastCounters = hc;
	ctx->lastDev = dev;
	ctx->directoryFetched = false;
	ctx->sortedRecords.clear();

	vxb_result_info& info = ctx->info;
	info.levels_total = (uint32_t)levels; info.levels_computed = (uint32_t)computed;
	info.block_count = hc.records; info.pad = 0;
	info.vertex_span = hc.vertices; info.index_span = hc.indices;
	info.trans_vertex_span = hc.transVertices; info.trans_index_span = hc.transIndices;
	info.vertex_total = hc.vertices; info.index_total = (uint64_t)hc.indices - 3ull * hc.degenerate;
	info.trans_vertex_total = hc.transVertices; info.trans_index_total = hc.transIndices;
	// statistics (TransVoxelImpl.cpp:528-531): BlocksCalculated counts every block of every computed level;
	// TrivialCells only those of processed (not skipped) blocks.  The statistics of a sharded run are the SUM over the
	// ranks (modulo 2^32): rank 0 carries the terms that do not depend on the work split.
	uint64_t blocksCalculated = 0, processedCells = (uint64_t)hc.nonSkippedLevel0 * 4096ull;
	for (int l = 0; l < computed; ++l)
	{
		const size_t b = region ? region->count[l] : blocksAtLevel(n, l);
		if (sharded && sh.rank != 0) continue;
		blocksCalculated += b;
		if (l) processedCells += b * 4096ull;
	}
	info.stats[0] = (uint32_t)blocksCalculated;
	info.stats[1] = (uint32_t)(processedCells - hc.nonTrivial);
	info.stats[2] = hc.nonTrivial;
	info.stats[3] = hc.degenerate;
	for (int i = 0; i < 16; ++i) info.stats[4 + i] = hc.perCase[i];
	for (int i = 0; i < 8; ++i) info.used_materials[i] = hc.usedMaterials[i];
	ctx->haveResult = true;
	if (!region && !sharded) { ctx->haveFullRun = (computed == levels); ctx->nextId = (uint32_t)totalBlocks; }
	else if (sharded) ctx->haveFullRun = false;
	return VXB_OK;
}

int vxb_polygonize(vxb_context* ctx, uint32_t maxLevels, uint32_t flags)
{
	return runPolygonize(ctx, maxLevels, flags, nullptr);
}

int vxb_polygonize_region(vxb_context* ctx, const float minCorner[3], const float maxCorner[3], uint32_t flags)
{
	if (!ctx || !minCorner || !maxCorner) return VXB_ERR_ARGUMENT;
	if (!ctx->haveGrid || !ctx->haveFullRun) return fail(ctx, VXB_ERR_STATE, "vxb_polygonize_region: needs a full vxb_polygonize of this grid first (the caches of that run are updated in place)");
	// GenerateBlockListForLevel, modification branch (TransVoxelImpl.cpp:429-465); corners are in OUTPUT (Y-up) coordinates
	Region region;
	vxb_region_info& ri = ctx->region;
	memset(&ri, 0, sizeof(ri));
	ri.levels = (uint32_t)ctx->levels;
	const float extent = (float)ctx->n;
	uint32_t id = ctx->nextId;
	for (int l = 0; l < ctx->levels; ++l)
	{
		const float blockMult = (float)((1 << l) * 16);
		float minBlock[3], maxBlock[3];
		for (int c = 0; c < 3; ++c)
		{
			float lo = floorf(minCorner[c] / blockMult - 1.0f) * blockMult, hi = floorf(maxCorner[c] / blockMult + 2.0f) * blockMult;
			lo = lo < 0.f ? 0.f : (lo > extent ? extent : lo);
			hi = hi < 0.f ? 0.f : (hi > extent ? extent : hi);
			ri.min_dirty[l][c] = lo; ri.max_dirty[l][c] = hi;
			minBlock[c] = lo / blockMult; maxBlock[c] = hi / blockMult;
		}
		// back to grid (Z-up) axes: x <- x, y <- output z, z <- output y  (:455-458)
		const unsigned lo3[3] = { (unsigned)minBlock[0], (unsigned)minBlock[2], (unsigned)minBlock[1] };
		const unsigned hi3[3] = { (unsigned)maxBlock[0], (unsigned)maxBlock[2], (unsigned)maxBlock[1] };
		size_t count = 1;
		for (int a = 0; a < 3; ++a)
		{
			region.rangeMin[l][a] = (int)lo3[a]; region.rangeMax[l][a] = (int)hi3[a];
			count *= hi3[a] > lo3[a] ? hi3[a] - lo3[a] : 0;
		}
		region.idStart[l] = id; region.count[l] = count;
		ri.id_start[l] = id; ri.block_count[l] = (uint32_t)count;
		id += (uint32_t)count;
	}
	const int r = runPolygonize(ctx, 0, flags, &region);
	if (r == VXB_OK) ctx->nextId = id;
	return r;
}

// ---- sharded runs: one grid, `world` ranks, one GPU each ----

int vxb_shard_configure(vxb_context* ctx, uint32_t rank, uint32_t world, uint32_t groupPlanes)
{
	if (!ctx) return VXB_ERR_ARGUMENT;
	if (!ctx->haveGrid) return fail(ctx, VXB_ERR_STATE, "vxb_shard_configure: no grid (vxb_cube_create or vxb_grid_set_device first)");
	const uint32_t n = ctx->n;
	if (world == 0 || world > 8 || rank >= world) return fail(ctx, VXB_ERR_ARGUMENT, "vxb_shard_configure: world in [1, 8], rank < world");
	if (groupPlanes == 0) groupPlanes = n / world;
	if (groupPlanes % 32 != 0 || n % groupPlanes != 0 || (n / groupPlanes) % world != 0)
		return fail(ctx, VXB_ERR_ARGUMENT, "vxb_shard_configure: group_planes must be a multiple of 32 planes that divides n into a multiple of `world` pieces");
	if (ctx->cube.active && (ctx->cube.groupPlanes != groupPlanes || ctx->cube.world != world || ctx->cube.rank != rank))
		return fail(ctx, VXB_ERR_ARGUMENT, "vxb_shard_configure: rank / world / group_planes differ from the cube's");
	const int coarseLo = coarseLoFor(n, ctx->levels);
	if (coarseLo >= ctx->levels) return fail(ctx, VXB_ERR_ARGUMENT, "vxb_shard_configure: the grid is too small to shard (n >= 64)");
	cudaSetDevice(ctx->device);
	releaseShard(ctx);
	vxb_context::Shard& sh = ctx->shard;
	sh.rank = rank; sh.world = world; sh.groupLayers = groupPlanes / 16; sh.sbLevel = coarseLo - 1;
	const size_t sb = blocksAtLevel(n, sh.sbLevel);
	VXB_CUDA(ctx, sh.sbMine.ensure(sb));
	VXB_CUDA(ctx, sh.sbWeight.ensure(sb));
	VXB_CUDA(ctx, sh.barrierBuf.ensure(8));
	VXB_CUDA(ctx, cudaMemset(sh.barrierBuf.p, 0, 64));
	sh.pagesBytes = sb * 4096 * sizeof(unsigned short);
	sh.validBytes = (sb + 255) & ~(size_t)255;
	{
		size_t offs[VXB_MAX_LEVELS];
		sh.latticeBytes = coarseLatticeLayout(n, ctx->levels, offs);
	}
	sh.latticeBytes = (sh.latticeBytes + 255) & ~(size_t)255;
	sh.bufBytes = sh.pagesBytes + sh.validBytes + sh.latticeBytes + 256; // pages | valid | lattices | barrier area (8 slots, epoch, CTA counter)
	const VmmApi& api = vmmApi();
	if (api.ok)
	{
		// exportable (the peers map it and store their pages into it)
		const CUmemAllocationProp prop = slabProp(ctx->device);
		size_t gran = 0;
		VXB_CU(ctx, api.granularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM));
		sh.bufBytes = (sh.bufBytes + gran - 1) / gran * gran;
		VXB_CU(ctx, api.addressReserve(&sh.pagesVa, sh.bufBytes, 0, 0, 0));
		VXB_CU(ctx, api.create(&sh.pagesHandle, sh.bufBytes, &prop, 0));
		std::vector<std::pair<CUdeviceptr, size_t> > scratch;
		const int r = mapRange(ctx, sh.pagesVa, sh.bufBytes, sh.pagesHandle, scratch);
		if (r != VXB_OK) return r;
		sh.pagesVmm = true;
		sh.pagesBuf = reinterpret_cast<unsigned char*>(sh.pagesVa);
	}
	else
	{
		VXB_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&sh.pagesBuf), sh.bufBytes));
	}
	VXB_CUDA(ctx, cudaMemset(sh.pagesBuf + sh.pagesBytes + sh.validBytes + sh.latticeBytes, 0, 256));
	sh.on = true;
	// virtual ranks over one shared upload: no shared even-lattice copy => level 1 gathers its tiles.  The coarse levels'
	// lattices move into the buffer the peers map (buildTensorMap -> buildCoarseLattices).
	if (!ctx->cube.active) ctx->latticeOff = true;
	{
		const int r = buildTensorMap(ctx);
		if (r != VXB_OK) return r;
	}
	if (ctx->graphExec) { cudaGraphExecDestroy(ctx->graphExec); ctx->graphExec = nullptr; }
	ctx->graphKey.clear();
	return VXB_OK;
}

int vxb_shard_buffers_get(vxb_context* ctx, vxb_shard_buffers* out)
{
	if (!ctx || !out) return VXB_ERR_ARGUMENT;
	const vxb_context::Shard& sh = ctx->shard;
	if (!sh.on) return fail(ctx, VXB_ERR_STATE, "vxb_shard_buffers_get: call vxb_shard_configure first");
	const size_t nb0 = ctx->n / 16, blocks0 = nb0 * nb0 * nb0;
	cudaSetDevice(ctx->device);
	VXB_CUDA(ctx, ctx->blockInfo.ensure(blocks0));
	memset(out, 0, sizeof(*out));
	out->block_info = ctx->blockInfo.p;
	out->block_info_bytes = blocks0;
	out->chunk_bytes = blocks0 / sh.world;
	out->pages = sh.pagesBuf;
	out->pages_bytes = sh.pagesBytes;
	out->valid = sh.pagesBuf + sh.pagesBytes;
	out->valid_bytes = sh.validBytes;
	out->super_level = (uint32_t)sh.sbLevel;
	return VXB_OK;
}

int vxb_shard_set_peer(vxb_context* ctx, uint32_t peer, void* pages, void* valid)
{
	if (!ctx) return VXB_ERR_ARGUMENT;
	vxb_context::Shard& sh = ctx->shard;
	if (!sh.on || peer >= sh.world || peer == sh.rank || !pages || !valid) return fail(ctx, VXB_ERR_ARGUMENT, "vxb_shard_set_peer: not configured, or bad peer / pointer");
	sh.peerPages[peer] = static_cast<unsigned short*>(pages);
	sh.peerValid[peer] = static_cast<unsigned char*>(valid);
	sh.peerLattice[peer] = static_cast<unsigned char*>(pages) + sh.pagesBytes + sh.validBytes; // one buffer: pages | valid | lattices
	sh.peerSet[peer] = true;
	return VXB_OK;
}

int vxb_shard_export(vxb_context* ctx, int* fd)
{
	if (!ctx || !fd) return VXB_ERR_ARGUMENT;
	const vxb_context::Shard& sh = ctx->shard;
	if (!sh.on || !sh.pagesVmm) return fail(ctx, VXB_ERR_STATE, "vxb_shard_export: needs vxb_shard_configure (and a driver with the virtual memory management API)");
	cudaSetDevice(ctx->device);
	int out = -1;
	VXB_CU(ctx, vmmApi().exportHandle(&out, sh.pagesHandle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
	*fd = out;
	return VXB_OK;
}

int vxb_shard_import(vxb_context* ctx, uint32_t peer, int fd)
{
	if (!ctx) return VXB_ERR_ARGUMENT;
	vxb_context::Shard& sh = ctx->shard;
	if (!sh.on || !sh.pagesVmm || peer >= sh.world || peer == sh.rank || fd < 0) return fail(ctx, VXB_ERR_ARGUMENT, "vxb_shard_import: not configured, or bad peer / descriptor");
	cudaSetDevice(ctx->device);
	const VmmApi& api = vmmApi();
	CUmemGenericAllocationHandle h = 0;
	VXB_CU(ctx, api.importHandle(&h, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
	sh.peerHandles.push_back(h);
	CUdeviceptr va = 0;
	VXB_CU(ctx, api.addressReserve(&va, sh.bufBytes, 0, 0, 0));
	const int r = mapRange(ctx, va, sh.bufBytes, h, sh.peerMapped);
	if (r != VXB_OK) return r;
	return vxb_shard_set_peer(ctx, peer, reinterpret_cast<void*>(va), reinterpret_cast<unsigned char*>(va) + sh.pagesBytes);
}

int vxb_shard_nccl_unique_id(vxb_nccl_id* id)
{
	if (!id) return VXB_ERR_ARGUMENT;
	const NcclApi& api = ncclApi();
	if (!api.ok) return fail(nullptr, VXB_ERR_CUDA, "vxb_shard_nccl_unique_id: libnccl.so.2 could not be loaded");
	const int r = api.getUniqueId(id);
	return r == 0 ? VXB_OK : failNccl(nullptr, "ncclGetUniqueId", r);
}

int vxb_shard_nccl_init(vxb_context* ctx, const vxb_nccl_id* id, uint32_t rank, uint32_t world)
{
	if (!ctx || !id) return VXB_ERR_ARGUMENT;
	const NcclApi& api = ncclApi();
	if (!api.ok) return fail(ctx, VXB_ERR_CUDA, "vxb_shard_nccl_init: libnccl.so.2 could not be loaded");
	cudaSetDevice(ctx->device);
	if (ctx->shard.comm) { api.commDestroy(ctx->shard.comm); ctx->shard.comm = nullptr; }
	const int r = api.commInitRank(&ctx->shard.comm, (int)world, *id, (int)rank);
	return r == 0 ? VXB_OK : failNccl(ctx, "ncclCommInitRank", r);
}

int vxb_polygonize_sharded(vxb_context* ctx, uint32_t phase, uint32_t flags)
{
	if (!ctx) return VXB_ERR_ARGUMENT;
	if (phase > 3) return fail(ctx, VXB_ERR_ARGUMENT, "vxb_polygonize_sharded: phase is 0, 1, 2 (the pieces around the two exchanges) or 3 (everything, NCCL inside)");
	return runPolygonize(ctx, 0, flags, nullptr, (int)phase);
}

int vxb_region_info_get(vxb_context* ctx, vxb_region_info* out)
{
	if (!ctx || !out) return VXB_ERR_ARGUMENT;
	*out = ctx->region;
	return VXB_OK;
}

int vxb_grid_update_blocks(vxb_context* ctx, uint32_t count, const uint32_t* blockCoords, const int8_t* dist, const uint8_t* mat, const uint8_t* blend)
{
	if (!ctx) return VXB_ERR_ARGUMENT;
	if (!ctx->haveGrid || !ctx->ownsGrid) return fail(ctx, VXB_ERR_STATE, "vxb_grid_update_blocks: needs a grid uploaded into context-owned storage");
	if (!count) return VXB_OK;
	if (!blockCoords || !dist) return fail(ctx, VXB_ERR_ARGUMENT, "vxb_grid_update_blocks: null pointer");
	cudaSetDevice(ctx->device);
	const size_t bytes = (size_t)count * 4096;
	VXB_CUDA(ctx, ctx->staging.ensure(bytes));
	VXB_CUDA(ctx, ctx->updCoords.ensure((size_t)count * 3));
	const uint32_t nb = ctx->n / 16;
	for (uint32_t i = 0; i < count * 3; ++i) if (blockCoords[i] >= nb) return fail(ctx, VXB_ERR_ARGUMENT, "vxb_grid_update_blocks: block coordinate out of range");
	VXB_CUDA(ctx, cudaMemcpyAsync(ctx->updCoords.p, blockCoords, (size_t)count * 12, cudaMemcpyHostToDevice, ctx->stream));
	const void* src[3] = { dist, mat, blend };
	uint8_t* dst[3] = { ctx->volDist.p, ctx->volMat.p, ctx->volBlend.p };
	for (int c = 0; c < 3; ++c)
	{
		if (!src[c]) continue; // channel unchanged
		VXB_CUDA(ctx, cudaMemcpyAsync(ctx->staging.p, src[c], bytes, cudaMemcpyHostToDevice, ctx->stream));
		vxb_unpack_block_list_kernel<<<count, VXB_THREADS, 0, ctx->stream>>>(reinterpret_cast<const uint4*>(ctx->staging.p), ctx->updCoords.p, dst[c], (int)ctx->n);
		VXB_CUDA(ctx, cudaGetLastError());
		VXB_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); // staging is reused by the next channel
	}
	ctx->haveResult = false;
	return VXB_OK;
}

int vxb_result_info_get(vxb_context* ctx, vxb_result_info* out)
{
	if (!ctx || !out) return VXB_ERR_ARGUMENT;
	if (!ctx->haveResult) return fail(ctx, VXB_ERR_STATE, "no result: call vxb_polygonize first");
	*out = ctx->info;
	return VXB_OK;
}

int vxb_result_download(vxb_context* ctx, vxb_block_record* records, void* vertices, uint32_t* indices, void* transVertices, uint32_t* transIndices)
{
	const int r = vxb_result_download_begin(ctx, records, vertices, indices, transVertices, transIndices);
	return r != VXB_OK ? r : vxb_result_download_end(ctx);
}

int vxb_result_download_begin(vxb_context* ctx, vxb_block_record* records, void* vertices, uint32_t* indices, void* transVertices, uint32_t* transIndices)
{
	if (!ctx) return VXB_ERR_ARGUMENT;
	if (!ctx->haveResult) return fail(ctx, VXB_ERR_STATE, "no result: call vxb_polygonize first");
	if (ctx->downloadOpen) return fail(ctx, VXB_ERR_STATE, "vxb_result_download_begin: the previous download was not ended");
	cudaSetDevice(ctx->device);
	const vxb_result_info& info = ctx->info;
	// directory first (small, into page-locked scratch), the arenas queued right behind it; the directory is sorted into the
	// reference's block order (level, then z,y,x = coord id; :395-401, :1278) on the host while the arenas are in flight
	const bool fetch = !ctx->directoryFetched;
	if (fetch && info.block_count)
	{
		if (!ctx->hostRecords.ensure(sizeof(vxb_block_record) * info.block_count)) return fail(ctx, VXB_ERR_CUDA, "vxb_result_download: pinned scratch allocation failed");
		VXB_CUDA(ctx, cudaMemcpyAsync(ctx->hostRecords.p, ctx->records.p, sizeof(vxb_block_record) * info.block_count, cudaMemcpyDeviceToHost, ctx->stream));
		VXB_CUDA(ctx, cudaEventRecord(ctx->evDir, ctx->stream));
	}
	if (vertices && info.vertex_span) VXB_CUDA(ctx, cudaMemcpyAsync(vertices, ctx->verts.p, info.vertex_span * sizeof(VxbVertex), cudaMemcpyDeviceToHost, ctx->stream));
	if (indices && info.index_span) VXB_CUDA(ctx, cudaMemcpyAsync(indices, ctx->idx.p, info.index_span * 4, cudaMemcpyDeviceToHost, ctx->stream));
	if (transVertices && info.trans_vertex_span) VXB_CUDA(ctx, cudaMemcpyAsync(transVertices, ctx->tverts.p, info.trans_vertex_span * sizeof(VxbVertex), cudaMemcpyDeviceToHost, ctx->stream));
	if (transIndices && info.trans_index_span) VXB_CUDA(ctx, cudaMemcpyAsync(transIndices, ctx->tidx.p, info.trans_index_span * 4, cudaMemcpyDeviceToHost, ctx->stream));
	if (fetch)
	{
		ctx->sortedRecords.resize(info.block_count);
		if (info.block_count)
		{
			VXB_CUDA(ctx, cudaEventSynchronize(ctx->evDir));
			memcpy(ctx->sortedRecords.data(), ctx->hostRecords.p, sizeof(vxb_block_record) * info.block_count);
		}
		std::sort(ctx->sortedRecords.begin(), ctx->sortedRecords.end(), [](const vxb_block_record& a, const vxb_block_record& b) {
			return a.level != b.level ? a.level < b.level : a.coord_id < b.coord_id; });
		ctx->directoryFetched = true;
	}
	if (records && info.block_count) memcpy(records, ctx->sortedRecords.data(), sizeof(vxb_block_record) * info.block_count);
	ctx->downloadOpen = true;
	ctx->downloadVerts = vertices;
	ctx->downloadTransVerts = transVertices;
	return VXB_OK;
}

int vxb_result_download_end(vxb_context* ctx)
{
	if (!ctx) return VXB_ERR_ARGUMENT;
	if (!ctx->downloadOpen) return fail(ctx, VXB_ERR_STATE, "vxb_result_download_end without vxb_result_download_begin");
	ctx->downloadOpen = false;
	cudaSetDevice(ctx->device);
	const vxb_result_info& info = ctx->info;
	void* const vertices = ctx->downloadVerts;
	void* const transVertices = ctx->downloadTransVerts;
	VXB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));

	// vertices of unmapped materials carry a marker (vxb_finish_vertex); restore the reference's all-zero textures
	ctx->unmapped.clear();
	bool anyUnmapped = false;
	for (unsigned i = 0; i < 256; ++i) if (!ctx->lutValid[i] && ((info.used_materials[i >> 5] >> (i & 31)) & 1u)) anyUnmapped = true;
	if (anyUnmapped)
	{
		auto fix = [&](void* base, uint64_t off, uint32_t count) {
			if (!base) return;
			VxbVertex* v = static_cast<VxbVertex*>(base) + off;
			for (uint32_t i = 0; i < count; ++i)
				if ((v[i].tex[0] & 0xFFu) == 0xFFu) { ctx->unmapped.push_back((uint8_t)(v[i].tex[0] >> 8)); v[i].tex[0] = 0; v[i].tex[1] = 0; }
		};
		for (const vxb_block_record& r : ctx->sortedRecords)
		{
			fix(vertices, r.vertex_offset, r.vertex_count);
			for (int f = 0; f < 6; ++f) fix(transVertices, r.trans_vertex_offset[f], r.trans_vertex_count[f]);
		}
	}
	return VXB_OK;
}

int vxb_result_device_arenas(vxb_context* ctx, const void** vertices, const uint32_t** indices, const void** transVertices, const uint32_t** transIndices)
{
	if (!ctx) return VXB_ERR_ARGUMENT;
	if (!ctx->haveResult) return fail(ctx, VXB_ERR_STATE, "no result: call vxb_polygonize first");
	if (vertices) *vertices = ctx->verts.p;
	if (indices) *indices = ctx->idx.p;
	if (transVertices) *transVertices = ctx->tverts.p;
	if (transIndices) *transIndices = ctx->tidx.p;
	return VXB_OK;
}

int vxb_result_select_lod(vxb_context* ctx, const float camera[3], float baseDistance, vxb_draw_lists* out)
{
	if (!ctx || !camera) return VXB_ERR_ARGUMENT;
	if (!ctx->haveResult) return fail(ctx, VXB_ERR_STATE, "no result: call vxb_polygonize first");
	cudaSetDevice(ctx->device);
	const uint32_t blocks = ctx->info.block_count;
	const size_t cap = blocks ? blocks : 1;
	VXB_CUDA(ctx, ctx->drawCmd.ensure(cap)); VXB_CUDA(ctx, ctx->drawInfo.ensure(cap));
	VXB_CUDA(ctx, ctx->drawTCmd.ensure(cap * 6)); VXB_CUDA(ctx, ctx->drawTInfo.ensure(cap * 6));
	VXB_CUDA(ctx, ctx->drawCounts.ensure(2));
	VXB_CUDA(ctx, cudaMemsetAsync(ctx->drawCounts.p, 0, 8, ctx->stream));
	VxbLodArgs a;
	for (int k = 0; k < 3; ++k) a.camera[k] = camera[k];
	a.baseDistance = baseDistance;
	a.counts = ctx->drawCounts.p;
	a.regular = ctx->drawCmd.p; a.regularInfo = ctx->drawInfo.p; a.transition = ctx->drawTCmd.p; a.transitionInfo = ctx->drawTInfo.p;
	a.capacity = (unsigned)(cap * 6);
	if (blocks) vxb_lod_select_kernel<<<(blocks + 255) / 256, 256, 0, ctx->stream>>>(ctx->lastDev, a, blocks);
	unsigned counts[2] = { 0, 0 };
	VXB_CUDA(ctx, cudaMemcpyAsync(counts, ctx->drawCounts.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
	VXB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	VXB_CUDA(ctx, cudaGetLastError());
	vxb_draw_lists& l = ctx->drawLists;
	l.regular_count = counts[0]; l.transition_count = counts[1];
	l.counts = ctx->drawCounts.p;
	l.regular = ctx->drawCmd.p; l.regular_info = ctx->drawInfo.p; l.transition = ctx->drawTCmd.p; l.transition_info = ctx->drawTInfo.p;
	if (out) *out = l;
	return VXB_OK;
}

int vxb_result_download_draws(vxb_context* ctx, vxb_draw_command* regular, vxb_draw_info* regularInfo, vxb_draw_command* transition, vxb_draw_info* transitionInfo)
{
	if (!ctx) return VXB_ERR_ARGUMENT;
	cudaSetDevice(ctx->device);
	const vxb_draw_lists& l = ctx->drawLists;
	if (regular && l.regular_count) VXB_CUDA(ctx, cudaMemcpy(regular, l.regular, sizeof(vxb_draw_command) * l.regular_count, cudaMemcpyDeviceToHost));
	if (regularInfo && l.regular_count) VXB_CUDA(ctx, cudaMemcpy(regularInfo, l.regular_info, sizeof(vxb_draw_info) * l.regular_count, cudaMemcpyDeviceToHost));
	if (transition && l.transition_count) VXB_CUDA(ctx, cudaMemcpy(transition, l.transition, sizeof(vxb_draw_command) * l.transition_count, cudaMemcpyDeviceToHost));
	if (transitionInfo && l.transition_count) VXB_CUDA(ctx, cudaMemcpy(transitionInfo, l.transition_info, sizeof(vxb_draw_info) * l.transition_count, cudaMemcpyDeviceToHost));
	return VXB_OK;
}

uint64_t vxb_result_unmapped_materials(vxb_context* ctx, uint8_t* ids, uint64_t capacity)
{
	if (!ctx) return 0;
	const uint64_t n = ctx->unmapped.size();
	if (ids) memcpy(ids, ctx->unmapped.data(), (size_t)std::min<uint64_t>(n, capacity));
	return n;
}

void* vxb_host_alloc(size_t bytes)
{
	void* p = nullptr;
	if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) return nullptr;
	return p;
}

void vxb_host_free(void* p) { if (p) cudaFreeHost(p); }

int vxb_kernel_ms(vxb_context* ctx, int which, float* ms, uint32_t* launches)
{
	if (!ctx || which < 0 || which > 9) return VXB_ERR_ARGUMENT;
	if (ms) *ms = ctx->kindMs[which];
	if (launches) *launches = ctx->kindLaunches[which];
	return VXB_OK;
}

} // extern "C"
