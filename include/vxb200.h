/* vxb200.h - C ABI of the B200-native Transvoxel polygonizer backend (libvxb200.so).
 *
 * This is the thin extern "C" layer between host code and the sm_100a CUDA kernels: plain pointers
 * and sizes only, no C++/torch types.  It replaces the reference's hot path
 *     Voxels::Polygonizer::Execute            /root/reference/src/TransVoxelImpl.cpp:74-79, :2153-2169
 *     TransVoxelRun::Execute (level loop)     /root/reference/src/TransVoxelImpl.cpp:468-538
 * and consumes exactly what that path reads from the grid store
 *     VoxelGrid::GetBlockData / GetMaterialBlockData / IsBlockEmpty   src/VoxelGrid.h:49-55
 *     (public form: Grid::GetBlockDistanceData / GetBlockMaterialData  include/Grid.h:127-139)
 * The C++ drop-in (voxels_b200/csrc/polygonizer_host.cpp: Voxels::Polygonizer, Modification,
 * PolygonSurface, BlockPolygons with the reference's vtable order) is built on top of these calls;
 * INTEGRATION.md shows the binding a maintainer of the reference would add.
 *
 * Conventions: grid coordinates are Z-up, x fastest; output vertices are Y-up exactly as the
 * reference emits them (TransVoxelImpl.cpp:1289-1291, :1336-1360).  All functions return 0 on
 * success or a negative vxb_status; vxb_last_error() gives the text.  No function falls back to
 * the CPU: without a CUDA device every call fails with VXB_ERR_CUDA.
 */
#ifndef VXB200_H
#define VXB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vxb_context vxb_context;

typedef enum vxb_status
{
	VXB_OK = 0,
	VXB_ERR_CUDA = -1,        /* CUDA runtime/driver error (text in vxb_last_error) */
	VXB_ERR_ARGUMENT = -2,    /* bad size (cube, power of two, >= 16), null pointer, ... */
	VXB_ERR_STATE = -3,       /* call order (no grid uploaded, no result yet, ...) */
	VXB_ERR_CAPACITY = -4     /* output arenas too small even after growing */
} vxb_status;

/* Flags of vxb_polygonize */
enum
{
	VXB_FLAG_NO_TRANSITIONS = 1,  /* do not emit transition meshes (BASELINE config 2); the reference always emits them */
	VXB_FLAG_KERNEL_TIMES = 2     /* bracket every kernel with CUDA events (vxb_kernel_ms); adds a little stream overhead */
};

/* One emitted block (a block that produced >= 1 regular vertex; reference: PushBlocksToResult
 * TransVoxelImpl.cpp:1266-1428).  Offsets index the arrays vxb_result_download fills. */
typedef struct vxb_block_record
{
	uint32_t level;
	uint32_t coord_id;            /* z*nb*nb + y*nb + x at that level (:399) */
	uint32_t id;                  /* Block id as the reference assigns it (:395-401) */
	uint32_t vertex_count;
	uint32_t index_count;         /* after the degenerate-triangle filter (:1300-1321) */
	uint32_t vertex_offset;       /* in vertices */
	uint32_t index_offset;        /* in indices */
	uint32_t trans_vertex_count[6];   /* face order = BlockPolygons::TransitionFaceId (include/Polygonizer.h:58-68) */
	uint32_t trans_index_count[6];
	uint32_t trans_vertex_offset[6];
	uint32_t trans_index_offset[6];
	uint32_t reserved;
} vxb_block_record;               /* 32 x 4 bytes */

typedef struct vxb_result_info
{
	uint32_t levels_total;        /* log2(n/16)+1 (:490) */
	uint32_t levels_computed;
	uint32_t block_count;         /* number of vxb_block_record */
	uint32_t pad;
	uint64_t vertex_span;         /* elements to allocate for each download array (spans include   */
	uint64_t index_span;          /* the gaps left by removed degenerate triangles)                 */
	uint64_t trans_vertex_span;
	uint64_t trans_index_span;
	uint64_t vertex_total;        /* exact totals over all records */
	uint64_t index_total;
	uint64_t trans_vertex_total;
	uint64_t trans_index_total;
	/* PolygonizationStatistics (include/Polygonizer.h:110-132): BlocksCalculated, TrivialCells,
	 * NonTrivialCells, DegenerateTrianglesRemoved, PerCaseCellsCount[16] */
	uint32_t stats[20];
	uint32_t used_materials[8];   /* bit per material id that occurs in an emitted vertex */
	float device_ms;              /* device time of the last vxb_polygonize (CUDA events) */
	uint32_t kernel_launches;     /* kernels launched by the last vxb_polygonize */
} vxb_result_info;

/* ---- context ------------------------------------------------------------------------------- */
int vxb_create(int device, vxb_context** out);
void vxb_destroy(vxb_context* ctx);
const char* vxb_last_error(const vxb_context* ctx);   /* ctx may be NULL: error of a failed vxb_create */
/* cudaStream_t the context launches on (as void*), for callers that time with their own events */
void* vxb_stream(vxb_context* ctx);
/* the context's second cudaStream_t (transition cells run on it next to the flat vertex / triangle kernels) */
void* vxb_exchange_stream(vxb_context* ctx);

/* ---- grid (input provider side: VoxelGrid accessors, src/VoxelGrid.h:49-55) ------------------ */
/* n^3 dense volumes in HOST memory, index (z*n + y)*n + x.  mat/blend may be NULL (zeros). */
int vxb_grid_upload_dense(vxb_context* ctx, uint32_t n, const int8_t* dist, const uint8_t* mat, const uint8_t* blend);
/* HOST memory, 16^3 blocks exactly as Grid::GetBlockDistanceData/GetBlockMaterialData return them,
 * concatenated in VoxelGrid block order (id = x + y*nb + z*nb*nb, VoxelGrid.h:139-144; 4096 B each). */
int vxb_grid_upload_blocks(vxb_context* ctx, uint32_t n, const int8_t* dist_blocks, const uint8_t* mat_blocks, const uint8_t* blend_blocks);
/* The grid in the reference's own serialised form, i.e. exactly the bytes Grid::PackForSave() returns
 * (VoxelGrid::PackForSave src/VoxelGrid.cpp:269-315: {version=1, w, d, h}, 3 x uint32 sizes per block, then per block
 * {flags, RLE distance, RLE material, RLE blend}; RLE = (length, value) byte pairs, VoxelGrid::CompressBlock :610-672).
 * HOST memory.  The blob is copied to the device as is (typically 10-30x smaller than the dense volumes) and the
 * run-length decoding (VoxelGrid::DecompressBlock :674-694) happens on the GPU. */
int vxb_grid_upload_packed(vxb_context* ctx, const void* blob, size_t size);
/* The same upload fed slab by slab, for a caller that is still gathering the grid store's blocks into `blob` (the drop-in's
 * Execute: VoxelGrid's per-block vectors -> one pinned blob).  At the call the 16-byte header and the COMPLETE size table
 * must be in place; the block data is produced on demand: before the library copies the data of the block layers
 * [layer0, layer1) (blocks layer0*nb*nb .. layer1*nb*nb-1 in VoxelGrid order) it calls produce(user, layer0, layer1) on the
 * calling thread, which must have written those blocks' {flags, distance, material, blend} bytes on return.  The copy and the
 * GPU decode of one slab run while the producer fills the next.  Layers are asked for in ascending order, each exactly once
 * (a context with a cube: only the layers of the pieces this rank backs).  produce = NULL: vxb_grid_upload_packed. */
typedef void (*vxb_pack_producer)(void* user, uint32_t layer0, uint32_t layer1);
int vxb_grid_upload_packed_streamed(vxb_context* ctx, const void* blob, size_t size, vxb_pack_producer produce, void* user);
/* Host-side helper (no GPU involved): writes the PackForSave form of dense n^3 volumes into `out` (capacity bytes).
 * Byte-identical to what the reference produces for the same voxels.  vxb_pack_dense_bound = worst-case size. */
size_t vxb_pack_dense_bound(uint32_t n);
int vxb_pack_dense(uint32_t n, const int8_t* dist, const uint8_t* mat, const uint8_t* blend, void* out, size_t capacity, size_t* written);

/* Dense volumes already resident in DEVICE memory (not copied; must stay valid; 16-byte aligned). */
int vxb_grid_set_device(vxb_context* ctx, uint32_t n, const int8_t* d_dist, const uint8_t* d_mat, const uint8_t* d_blend);
/* Device pointers of the context-owned dense volumes (after an upload), for tools/benchmarks. */
int vxb_grid_device_pointers(vxb_context* ctx, const int8_t** d_dist, const uint8_t** d_mat, const uint8_t** d_blend);

/* ---- device-resident grid store (the steps before the polygonizer; reference src/VoxelGrid.cpp) ------------------
 * A built-in procedural surface: the device-side analogue of a client's Voxels::VoxelSurface (include/VoxelSurface.h:35-40;
 * a client callback cannot run on the GPU, these can).  voxels_b200/csrc/vxb_surfaces.h evaluates them with bit-identical
 * floats on the device and on the host (the test harness serves the same functions to the unmodified reference grid store).
 *   SPHERE   p = {cx, cy, cz, r}: d = |x - c| - r             material / blend: constants
 *   PLANE    p = {nx, ny, nz, d0}: d = n . x - d0              material / blend: constants
 *   TERRAIN  p = {N, ox, oy}: the seeded Perlin terrain of BASELINE configs[1..4]: heightfield 0.5 N +
 *            0.18 N fbm2 + 6 perlin3 detail, 3 material bands + ore pockets, smooth blends; N = horizontal period (the grid
 *            edge), (ox, oy) = window origin; seed = permutation table
 * Distances are clamped to +-100 before quantisation (keeps the reference's char conversion defined, VoxelGrid.cpp:37-40). */
enum { VXB_SURFACE_SPHERE = 0, VXB_SURFACE_PLANE = 1, VXB_SURFACE_TERRAIN = 2 };
typedef struct vxb_surface
{
	uint32_t kind;
	uint32_t material, blend;     /* SPHERE / PLANE */
	uint32_t seed;                /* TERRAIN */
	float p[8];
} vxb_surface;
/* Grid::Create(n, n, n, start_x, start_y, start_z, step, &surface) (VoxelGrid.cpp:79-132) into context-owned dense volumes -
 * or, when the context has a cube (sharded runs), into the pieces this rank backs: sample, round away from zero, clamp to +-4. */
int vxb_grid_fill(vxb_context* ctx, uint32_t n, const vxb_surface* surface, const float start[3], float step);
/* VoxelGrid::InjectSurface (VoxelGrid.cpp:388-488) on the device copy: position / extents in grid (Z-up) coordinates, type =
 * InjectionType (0 IT_Add, 1 IT_SubtractAddInner, 2 IT_Subtract; include/Grid.h).  The surface is sampled relative to
 * `position` exactly as the reference samples the client's surface (:419-428).  out_min / out_max = the modified box the
 * reference returns (Y-up "DX style", :479-487) = what Modification::MinCornerModified / MaxCornerModified take.
 * Needs a context-owned grid (upload or fill). */
int vxb_grid_inject_surface(vxb_context* ctx, const float position[3], const float extents[3], const vxb_surface* surface, int type,
	float out_min[3], float out_max[3]);
/* VoxelGrid::InjectMaterial (VoxelGrid.cpp:490-584). */
int vxb_grid_inject_material(vxb_context* ctx, const float position[3], const float extents[3], uint32_t material, int add_subtract_blend,
	float out_min[3], float out_max[3]);
/* Grid::PackForSave of the device grid (VoxelGrid::CompressBlock :610-672 + PackForSave :269-315): every block run-length coded on
 * the GPU (raw fallback when RLE is ineffective, BF_Empty flag), then the blob copied to HOST memory `out` (capacity bytes,
 * vxb_pack_dense_bound is enough).  Byte-identical to what the reference produces for the same voxels. */
int vxb_grid_pack(vxb_context* ctx, void* out, size_t capacity, size_t* written);
/* The dense volumes back to HOST memory (n^3 bytes each; any pointer may be NULL). */
int vxb_grid_download_dense(vxb_context* ctx, int8_t* dist, uint8_t* mat, uint8_t* blend);

/* MaterialMap::GetMaterial pre-tabulated (include/MaterialMap.h:19-30): table = 256 x
 * {DiffuseIds0[3], DiffuseIds1[3]}, valid = 256 bytes (0 = GetMaterial returned nullptr).
 * NULL table = identity map, NULL valid = all valid. */
int vxb_set_materials(vxb_context* ctx, const uint8_t* table, const uint8_t* valid);

/* ---- polygonization -------------------------------------------------------------------------- */
/* Full run over LOD levels [0, max_levels) (0 = all, as the reference).  Results stay in device
 * arenas until the next call.  Blocks until the device work is done. */
int vxb_polygonize(vxb_context* ctx, uint32_t max_levels, uint32_t flags);
int vxb_result_info_get(vxb_context* ctx, vxb_result_info* out);

/* ---- incremental re-polygonization (Execute with a Modification, TransVoxelImpl.cpp:362-364, :429-465) ---- */
/* Replaces `count` 16^3 blocks of the context-owned grid (after an edit such as Grid::InjectSurface).
 * block_coords = count x {x, y, z} block coordinates; dist/mat/blend = count x 4096 bytes each in the same order
 * (Grid::GetBlockDistanceData / GetBlockMaterialData layout); mat and blend may be NULL (unchanged). HOST memory. */
int vxb_grid_update_blocks(vxb_context* ctx, uint32_t count, const uint32_t* block_coords, const int8_t* dist, const uint8_t* mat, const uint8_t* blend);

/* Per level: the dirty box of the last vxb_polygonize_region and the ids it handed out.  A caller that keeps a
 * surface erases its blocks whose minimal corner lies in [min_dirty, max_dirty) (:443-450) and appends the new ones;
 * ids id_start .. id_start+block_count-1 are Modification::GetModifiedBlocks (:463). Corners are OUTPUT (Y-up) coordinates. */
typedef struct vxb_region_info
{
	uint32_t levels;
	uint32_t pad;
	float min_dirty[12][3];
	float max_dirty[12][3];
	uint32_t id_start[12];
	uint32_t block_count[12];
} vxb_region_info;

/* Re-polygonizes the blocks of every level around the box [min_corner, max_corner] (Modification::MinCornerModified /
 * MaxCornerModified, OUTPUT coordinates: the box Grid::InjectSurface returns).  Needs a preceding full vxb_polygonize of
 * the same context: its consistency / material caches are updated in place exactly like the reference's PolygonMap
 * caches (bits are only ever set, votes only overwrite), block ids continue the context's running counter.  The
 * result (vxb_result_info / vxb_result_download) holds ONLY the re-created blocks. */
int vxb_polygonize_region(vxb_context* ctx, const float min_corner[3], const float max_corner[3], uint32_t flags);
int vxb_region_info_get(vxb_context* ctx, vxb_region_info* out);

/* ---- Sharded runs: ONE grid polygonized by `world` ranks, one GPU each (BASELINE configs[3]) ----
 * The reference balances its OpenMP block loop by blocks (TransVoxelImpl.cpp:500-503); so does this: WORK is dealt by
 * blocks, independently of where the DATA lives.
 *
 * Data.  The cube (vxb_cube_*): each volume is ONE contiguous n^3 virtual range on every rank; the z-axis is cut into
 * pieces of `group_planes` planes, piece p backed by the HBM of rank p % world (cyclic, so that a terrain's surface - a
 * few z-layers - is spread over all ranks' memories) and mapped into every peer over NVLink.  Kernels and TMA tensor maps
 * address the cube exactly like a single-GPU grid.  A fourth channel holds the even-lattice copy the level-1 tiles read.
 *
 * Work.  Every rank streams its OWN pieces once (vxb_scan_kernel: flags per 16^3 block); ONE ncclAllGather makes the
 * 1-byte-per-block info complete everywhere (exchange 0).  From it every rank derives the same work split: the
 * "super-blocks" (blocks of the last level with more than 4096 blocks) in coordinate order, cut where the running count of
 * surface-crossing level-0 blocks passes total * r / world.  A rank polygonizes every level up to the super-block
 * level inside its super-blocks - the material votes of a cell only read its own children (:763-837), so this chain is
 * rank-local - and stores the pages of its super-blocks into every peer's page buffer (peer stores, exchange 1, ordered by a
 * second tiny all-gather).  The few coarse levels above are classified by every rank (identical votes), and block b of
 * level l is emitted by rank (b + l) % world.  Block ids are the reference's full-run ids (:395-401), so the directories
 * of the ranks concatenate and sort into exactly the single-GPU directory; statistics are the sum over the ranks.
 *
 *   vxb_cube_create(ctx, n, rank, world, group_planes)      reserve the ranges, back + map the local pieces (0 = n / world)
 *   vxb_cube_export / vxb_cube_import                        one POSIX file descriptor per (channel, piece), sent to the peers
 *   vxb_shard_configure(ctx, rank, world, group_planes)     work-split state + the page buffer (needs a grid: the cube, or for
 *                                                            virtual ranks of one process a shared vxb_grid_set_device)
 *   vxb_shard_export / vxb_shard_import                      the page buffer, likewise (or vxb_shard_set_peer with raw pointers)
 *   vxb_shard_nccl_unique_id / vxb_shard_nccl_init           the communicator of the two all-gathers (libnccl.so.2 via dlopen; the
 *                                                            id travels by the caller's own means, e.g. torch.distributed)
 *   vxb_polygonize_sharded(ctx, 3, flags)                    the whole step, one stream-ordered sequence (replayed as a CUDA graph)
 *   vxb_polygonize_sharded(ctx, 0 | 1 | 2, flags)            the three pieces around the two exchanges, for callers that do
 *                                                            the exchanges themselves (tests: device copies between virtual ranks):
 *                                                            after 0 gather block_info (rank r's chunk = bytes [r, r+1) * chunk_bytes),
 *                                                            after 1 synchronise all ranks; 2 blocks and delivers the result.
 * VXB_ERR_CAPACITY means the arenas were grown: every rank repeats the run. */
typedef struct vxb_nccl_id { char internal[128]; } vxb_nccl_id;   /* = ncclUniqueId */
typedef struct vxb_shard_buffers
{
	void* block_info;             /* device: one byte per level-0 block, rank-major (the layout of the all-gather) */
	uint64_t block_info_bytes;
	uint64_t chunk_bytes;         /* block_info_bytes / world */
	void* pages;                  /* device: {material id, blend} of every cell of every super-block, block-major (z,y,x) */
	uint64_t pages_bytes;
	void* valid;                  /* device: one byte per super-block */
	uint64_t valid_bytes;
	uint32_t super_level;
	uint32_t pad;
} vxb_shard_buffers;
int vxb_shard_configure(vxb_context* ctx, uint32_t rank, uint32_t world, uint32_t group_planes);
int vxb_shard_buffers_get(vxb_context* ctx, vxb_shard_buffers* out);
int vxb_shard_set_peer(vxb_context* ctx, uint32_t peer, void* pages, void* valid);
int vxb_shard_export(vxb_context* ctx, int* fd);
int vxb_shard_import(vxb_context* ctx, uint32_t peer, int fd);
int vxb_shard_nccl_unique_id(vxb_nccl_id* id);
int vxb_shard_nccl_init(vxb_context* ctx, const vxb_nccl_id* id, uint32_t rank, uint32_t world);
int vxb_polygonize_sharded(vxb_context* ctx, uint32_t phase, uint32_t flags);

/* The cube: see above.  Every piece (group_planes * n^2 bytes; the lattice channel an eighth of that) must be a multiple of
 * the allocation granularity (2 MiB); when only the lattice pieces are not, the cube has three channels and level 1 gathers
 * its tiles.  vxb_cube_info: number of pieces, channels and bytes per piece of each channel (uint64_t[4]).
 * vxb_cube_piece: device pointers of piece `piece` (any rank's: the whole cube is addressable) and its size in bytes. */
int vxb_cube_create(vxb_context* ctx, uint32_t n, uint32_t rank, uint32_t world, uint32_t group_planes);
int vxb_cube_info(vxb_context* ctx, uint32_t* pieces, uint32_t* channels, uint64_t* piece_bytes);
int vxb_cube_export(vxb_context* ctx, uint32_t channel, uint32_t piece, int* fd);
int vxb_cube_import(vxb_context* ctx, uint32_t channel, uint32_t piece, int fd);
int vxb_cube_piece(vxb_context* ctx, uint32_t piece, int8_t** dist, uint8_t** mat, uint8_t** blend, uint64_t* bytes_per_channel);
/* Copies the directory (sorted by level, then coord_id = the reference's block order) and the
 * arenas to HOST memory sized from vxb_result_info spans.  Vertices are 48-byte
 * Voxels::PolygonVertex (include/Polygonizer.h:14-48).  Any pointer may be NULL to skip it. */
int vxb_result_download(vxb_context* ctx, vxb_block_record* records, void* vertices, uint32_t* indices,
	void* trans_vertices, uint32_t* trans_indices);
/* The same download in two steps: _begin queues the copies of the arenas and returns as soon as the directory has arrived
 * and been sorted (`records` is filled), so that the caller can build its per-block views (pointers into the arenas:
 * PushBlocksToResult, src/TransVoxelImpl.cpp:1274-1293) while the arenas are still in flight; _end waits for them (and
 * restores the all-zero textures of unmapped materials).  The arenas must not be read, and no run started, in between. */
int vxb_result_download_begin(vxb_context* ctx, vxb_block_record* records, void* vertices, uint32_t* indices,
                              void* trans_vertices, uint32_t* trans_indices);
int vxb_result_download_end(vxb_context* ctx);

/* ---- consumer side: the result as indirect draws, without leaving HBM (doc_source/Rendering.md:18-58) ------------------
 * The arenas are merged vertex / index buffers: a block is drawn with {index_count, first_index, base_vertex} (indices are
 * block-local).  vxb_result_device_arenas hands the device pointers to the renderer (CUDA-graphics interop / external
 * memory).  vxb_result_select_lod runs, per frame, the cut through the LOD octree on the GPU: node of level l around the
 * camera is refined while the camera is closer than base_distance * 2^l to its centre; for every drawn block of level > 0
 * `block_adj` has bit f set when the neighbour across face f (BlockPolygons::TransitionFaceId order) is drawn finer - then
 * the block's transition mesh f is drawn too and the vertex shader selects the secondary position of the vertices whose
 * mask (SecondaryPosition.w) is covered by block_adj (Rendering.md:44-52).  Both lists are VkDrawIndexedIndirectCommand /
 * D3D12_DRAW_INDEXED_ARGUMENTS arrays with first_instance = the draw's index into the info array. */
typedef struct vxb_draw_command { uint32_t index_count, instance_count, first_index; int32_t base_vertex; uint32_t first_instance; } vxb_draw_command;
typedef struct vxb_draw_info { uint32_t block_id, level, block_adj, face; /* face = 0xFFFFFFFF: the regular mesh */ } vxb_draw_info;
typedef struct vxb_draw_lists
{
	uint32_t regular_count, transition_count;   /* also resident on the device: counts[0], counts[1] (for *IndirectCount draws) */
	const uint32_t* counts;                     /* device */
	const vxb_draw_command* regular; const vxb_draw_info* regular_info;          /* device, regular_count entries */
	const vxb_draw_command* transition; const vxb_draw_info* transition_info;    /* device, transition_count entries */
} vxb_draw_lists;
int vxb_result_device_arenas(vxb_context* ctx, const void** vertices, const uint32_t** indices, const void** trans_vertices, const uint32_t** trans_indices);
int vxb_result_select_lod(vxb_context* ctx, const float camera[3], float base_distance, vxb_draw_lists* out);
/* the two lists copied to HOST memory (tests, debugging); capacities in entries, any pointer may be NULL */
int vxb_result_download_draws(vxb_context* ctx, vxb_draw_command* regular, vxb_draw_info* regular_info, vxb_draw_command* transition, vxb_draw_info* transition_info);

/* GetMaterial(id) == nullptr handling (TransVoxelImpl.cpp:1364-1368: textures stay zero, one LS_Error log per
 * vertex): after vxb_result_download the material ids of those vertices, in the reference's logging order
 * (level, block, regular vertices, then transition faces), can be read here.  Returns the count; copies at
 * most `capacity` ids. */
uint64_t vxb_result_unmapped_materials(vxb_context* ctx, uint8_t* ids, uint64_t capacity);

/* Page-locked host memory for the upload/download buffers (plain cudaHostAlloc/cudaFreeHost). */
void* vxb_host_alloc(size_t bytes);
void vxb_host_free(void* p);

/* Initial arena capacities in elements (0 = keep default).  Arenas grow and the run repeats
 * automatically on overflow; this only avoids the retry. */
int vxb_set_capacity(vxb_context* ctx, uint64_t vertices, uint64_t indices, uint64_t trans_vertices, uint64_t trans_indices);

/* Per-kernel device time of the last vxb_polygonize run with VXB_FLAG_KERNEL_TIMES, for bench.py's roofline line:
 * which = 0: vxb_scan_kernel (streams the level-0 distance volume once), 1: block info, sign-mix pyramid + selection kernels,
 * 2: vxb_block_kernel<1>/<2> (levels >= 1: classification, votes, decisions) + vxb_decide_kernel<4096>, 3: vxb_block_kernel<0>
 * (level 0: classification, decisions, vertices and triangles in one pass), 4: vxb_vertex_kernel (levels >= 1),
 * 5: vxb_triangle_kernel (levels >= 1), 6: vxb_transition_kernel + vertices, 7: vxb_finish_kernel; sharded runs (phase 3) also
 * 8: exchange 0 (lattice publication + ncclAllGather of the block info), 9: exchange 1 (page publication + the ordering all-gather).
 * Milliseconds, summed per kind. */
int vxb_kernel_ms(vxb_context* ctx, int which, float* ms, uint32_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* VXB200_H */
