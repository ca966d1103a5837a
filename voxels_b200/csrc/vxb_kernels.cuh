// sm_100a kernels of the Transvoxel polygonizer.  Included by vxb200.cu only.
//
// Pipeline per vxb_polygonize call (reference: TransVoxelRun::Execute, TransVoxelImpl.cpp:468-538):
//   vxb_scan_kernel          streams the level-0 distance volume ONCE (coalesced 16-byte loads) and
//                            reduces, per 16^3 block, sign/zero/lattice flags + a run-count bound
//                            (VoxelGrid::CompressBlock's BF_Empty rule, VoxelGrid.cpp:610-672).
//   vxb_block_info_kernel    turns the raw flags into BF_Empty / sign-mix bits (exact slow path when
//                            the run-count bound is inconclusive).
//   per LOD level:
//     vxb_select_kernel      block walk (GenerateBlockListForLevel :385-466 + AreBlockAndNeighborsEmpty
//                            :1511-1527): appends the blocks that can produce output or side effects.
//     vxb_polygonize_kernel  persistent CTAs, one 16^3 block at a time: TMA-staged 17^3 sample tile,
//                            case codes, ordered (ballot + scan) compaction that reproduces the
//                            reference's serial vertex numbering, vertex/triangle emission, degenerate
//                            filter, transition cells (PolygonizeBlock :1529-1750,
//                            GenerateTransitionCells :1754-2131, PushBlocksToResult :1266-1428).
#pragma once

#define VXB_MAX_LEVELS 12
#define VXB_THREADS 256
#define VXB_TILE_PITCH 32
#define VXB_TILE_BYTES (17 * 17 * VXB_TILE_PITCH)

// blockInfo bits (per level-0 block)
#define VXB_BI_NEG 1u
#define VXB_BI_NONNEG 2u
#define VXB_BI_EMPTY 4u
#define VXB_BI_NEG_E 8u      // level-1 lattice (even coordinates)
#define VXB_BI_NONNEG_E 16u

struct VxbCounters
{
	unsigned int vertices, indices, transVertices, transIndices; // arena cursors (may exceed capacity = overflow)
	unsigned int records;
	unsigned int nonSkippedLevel0;
	unsigned int nonTrivial;
	unsigned int degenerate;
	unsigned int perCase[16];
	unsigned int usedMaterials[8];
	unsigned int workCount[VXB_MAX_LEVELS];
	unsigned int workCursor[VXB_MAX_LEVELS];
	unsigned int emitCount[VXB_MAX_LEVELS]; // blocks with non-trivial cells, per level (vxb_classify_kernel)
	unsigned int emitCursor;
	unsigned int bigCount, bigCursor;       // rejected by the small emit tier
	unsigned int genCount, genCursor;       // rejected by the big emit tier -> generic kernel
};

struct VxbDev
{
	VxbGrid grid;
	int n, levels, lastLevel;
	unsigned int* scanFlags;
	unsigned char* blockInfo;
	unsigned int* consPages;                    // [nb0^3][128]
	unsigned char* consValid;                   // [nb0^3]
	unsigned short* cachePages[VXB_MAX_LEVELS]; // level l >= 1: [nb_l^3][4096] {id, blend}
	unsigned char* cacheValid[VXB_MAX_LEVELS];
	unsigned int* worklist;
	unsigned int workBase[VXB_MAX_LEVELS];
	unsigned int idBase[VXB_MAX_LEVELS];
	VxbVertex* verts; unsigned int* idx; VxbVertex* tverts; unsigned int* tidx;
	vxb_block_record* records;
	unsigned int vcap, icap, tvcap, ticap, rcap;
	VxbCounters* counters;
	const VxbMaterialLut* lut;
	int transitions;
	unsigned int* emitList;   // [workBase[l] + i] = level<<28 | coordId
	unsigned int* ntScratch;  // 128 words of non-trivial bits per emit-list entry
	unsigned int* bigList;    // emit-list indices
	unsigned int* genList;
};

// ------------------------------------------------------------------------------------------------
// small PTX wrappers (mbarrier + TMA), sm_90+/sm_100a
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned vxb_smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void vxb_mbar_init(unsigned long long* bar, unsigned count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(vxb_smem_addr(bar)), "r"(count));
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void vxb_mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(vxb_smem_addr(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void vxb_mbar_wait(unsigned long long* bar, unsigned parity)
{
	const unsigned addr = vxb_smem_addr(bar);
	unsigned done, spins = 0;
	do
	{
		asm volatile(
			"{\n"
			".reg .pred p;\n"
			"mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
			"selp.u32 %0, 1, 0, p;\n"
			"}\n" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
		if (!done && ++spins > (1u << 26)) __trap(); // a TMA that never lands must fail the launch, not hang the GPU
	} while (!done);
}

// 3-D tiled TMA load global -> shared, completion on an mbarrier (SASS: UTMALDG)
__device__ __forceinline__ void vxb_tma_load_3d(void* dst, const CUtensorMap* map, int x, int y, int z, unsigned long long* bar)
{
	asm volatile(
		"cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
		::"r"(vxb_smem_addr(dst)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(vxb_smem_addr(bar)) : "memory");
}

__device__ __forceinline__ void vxb_fence_proxy_async()
{
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// exclusive scan over the 256 threads of the CTA; every thread gets the grand total too
__device__ __forceinline__ unsigned vxb_block_scan(unsigned v, unsigned* warpSums /*[8]*/, unsigned& total)
{
	const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	unsigned inc = v;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1)
	{
		const unsigned t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
		if (lane >= (unsigned)o) inc += t;
	}
	__syncthreads(); // previous users of warpSums are done
	if (lane == 31) warpSums[warp] = inc;
	__syncthreads();
	unsigned base = 0, tot = 0;
#pragma unroll
	for (int w = 0; w < VXB_THREADS / 32; ++w)
	{
		const unsigned s = warpSums[w];
		if ((unsigned)w < warp) base += s;
		tot += s;
	}
	total = tot;
	return base + inc - v;
}

// ------------------------------------------------------------------------------------------------
// K1: scan - one pass over the level-0 distance volume
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned vxb_zero_bytes(unsigned w) // 0x80 in every byte of w that is zero (exact)
{
	return ~(((w & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | w) & 0x80808080u;
}

__global__ void __launch_bounds__(VXB_THREADS) vxb_scan_kernel(const signed char* __restrict__ dist, int n, unsigned int* __restrict__ scanFlags)
{
	__shared__ unsigned sFlags[8];
	__shared__ unsigned sChanges[8];
	const int nb = n >> 4;
	const int bx0 = blockIdx.x * 8, by = blockIdx.y, bz = blockIdx.z;
	const int tid = threadIdx.x;
	const int j = tid & 7, rg = tid >> 3;
	if (tid < 8) { sFlags[tid] = 0; sChanges[tid] = 0; }
	__syncthreads();

	unsigned neg = 0, pos = 0, zero = 0, negE = 0, nonnegE = 0, changes = 0;
	if (bx0 + j < nb)
	{
		const uint4* src = reinterpret_cast<const uint4*>(dist + (((size_t)bz * 16) * n + (size_t)by * 16) * n + (size_t)(bx0 + j) * 16);
		uint4 rows[8];
#pragma unroll
		for (int i = 0; i < 8; ++i)
		{
			const int row = i * 32 + rg, y = row & 15, z = row >> 4;
			rows[i] = __ldg(src + (((size_t)z * n + y) * n >> 4));
		}
#pragma unroll
		for (int i = 0; i < 8; ++i)
		{
			const int row = i * 32 + rg, y = row & 15, z = row >> 4;
			const unsigned w[4] = { rows[i].x, rows[i].y, rows[i].z, rows[i].w };
			const bool even = !((y | z) & 1);
#pragma unroll
			for (int q = 0; q < 4; ++q)
			{
				const unsigned zb = vxb_zero_bytes(w[q]);
				neg |= w[q] & 0x80808080u;
				zero |= zb;
				pos |= ~w[q] & 0x80808080u & ~zb;
				if (even) { negE |= w[q] & 0x00800080u; nonnegE |= ~w[q] & 0x00800080u; }
				// adjacent-byte changes inside the 16-byte row: compare byte k with byte k+1
				const unsigned nxt = (q < 3) ? w[q + 1] : (w[3] >> 24);
				const unsigned shifted = (w[q] >> 8) | (nxt << 24);
				const unsigned x = w[q] ^ shifted;
				unsigned diff = ((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x; // bit 7 of each byte set iff byte != 0
				diff &= (q < 3) ? 0x80808080u : 0x00808080u;               // byte 15 has no right neighbour in the row
				changes += __popc(diff);
			}
		}
	}
	unsigned f = (neg ? 1u : 0u) | (pos ? 2u : 0u) | (zero ? 4u : 0u) | (negE ? 8u : 0u) | (nonnegE ? 16u : 0u);
	f |= __shfl_xor_sync(0xFFFFFFFFu, f, 8); f |= __shfl_xor_sync(0xFFFFFFFFu, f, 16);
	changes += __shfl_xor_sync(0xFFFFFFFFu, changes, 8); changes += __shfl_xor_sync(0xFFFFFFFFu, changes, 16);
	if ((tid & 31) < 8) { atomicOr(&sFlags[j], f); atomicAdd(&sChanges[j], changes); }
	__syncthreads();
	if (tid < 8 && bx0 + tid < nb)
		scanFlags[((size_t)bz * nb + by) * nb + bx0 + tid] = sFlags[tid] | (sChanges[tid] << 8);
}

// K1b: raw flags -> blockInfo.  One thread per level-0 block.
__global__ void vxb_block_info_kernel(const signed char* __restrict__ dist, int n, const unsigned int* __restrict__ scanFlags, unsigned char* __restrict__ blockInfo)
{
	const int nb = n >> 4;
	const size_t total = (size_t)nb * nb * nb;
	for (size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < total; b += (size_t)gridDim.x * blockDim.x)
	{
		const unsigned f = scanFlags[b];
		const bool neg = f & 1, pos = f & 2, zero = f & 4;
		const unsigned changes = f >> 8; // value changes inside 16-byte rows (lower bound of run ends)
		bool empty = (neg != pos) && !zero; // every sample non-zero with one strict sign
		if (empty)
		{
			// CompressBlock stores the block raw (and never flags it empty) once 2048 run ends were seen.
			// runs <= changes + 255 (row seams) + 16 (255-length splits) + 1
			if (changes >= 2048u) empty = false;
			else if (changes + 272u >= 2049u)
			{
				// inconclusive: walk the block exactly (rare: strictly one-signed AND very noisy)
				const int bx = (int)(b % nb), by = (int)((b / nb) % nb), bz = (int)(b / ((size_t)nb * nb));
				unsigned counter = 0, runEnds = 0;
				int last = 0;
				for (int i = 0; i < 4096 && empty; ++i)
				{
					const int x = i & 15, y = (i >> 4) & 15, z = i >> 8;
					const int cur = dist[(((size_t)bz * 16 + z) * n + (size_t)by * 16 + y) * n + (size_t)bx * 16 + x];
					if (i == 0) last = cur;
					if (last == cur && counter < 0xFF) { ++counter; continue; }
					++runEnds; counter = 1; last = cur;
					if (1 + 2 * runEnds > 4096) empty = false;
				}
			}
		}
		blockInfo[b] = (unsigned char)((neg ? VXB_BI_NEG : 0) | ((pos || zero) ? VXB_BI_NONNEG : 0) | (empty ? VXB_BI_EMPTY : 0)
			| ((f & 8) ? VXB_BI_NEG_E : 0) | ((f & 16) ? VXB_BI_NONNEG_E : 0));
	}
}

// Upload helper: 16^3 blocks (VoxelGrid block order) -> dense volume.  One CTA per block, 256 threads x 16 bytes.
__global__ void __launch_bounds__(VXB_THREADS) vxb_unpack_blocks_kernel(const uint4* __restrict__ blocks, unsigned char* __restrict__ dense, int n)
{
	const int nb = n >> 4;
	const size_t b = blockIdx.x;
	const int bx = (int)(b % nb), by = (int)((b / nb) % nb), bz = (int)(b / ((size_t)nb * nb));
	const int row = threadIdx.x, y = row & 15, z = row >> 4;
	const uint4 v = blocks[b * 256 + row];
	*reinterpret_cast<uint4*>(dense + (((size_t)bz * 16 + z) * n + (size_t)by * 16 + y) * n + (size_t)bx * 16) = v;
}

// ------------------------------------------------------------------------------------------------
// K2: block selection for one level
// ------------------------------------------------------------------------------------------------
__global__ void vxb_select_kernel(VxbDev d, int level)
{
	const int m = 1 << level, nb = d.n / 16 / m, nb0 = d.n / 16;
	const unsigned total = (unsigned)nb * nb * nb;
	for (unsigned base = blockIdx.x * blockDim.x; base < total; base += gridDim.x * blockDim.x)
	{
		const unsigned b = base + threadIdx.x;
		bool take = false, counted = false;
		if (b < total)
		{
			const int bx = b % nb, by = (b / nb) % nb, bz = b / (nb * nb);
			if (level == 0)
			{
				// AreBlockAndNeighborsEmpty :1511-1527
				bool skip = true;
				for (int z = -1; z < 2 && skip; ++z) for (int y = -1; y < 2 && skip; ++y) for (int x = -1; x < 2; ++x)
				{
					const int cx = min(max(bx + x, 0), nb - 1), cy = min(max(by + y, 0), nb - 1), cz = min(max(bz + z, 0), nb - 1);
					if (!(d.blockInfo[((size_t)cz * nb + cy) * nb + cx] & VXB_BI_EMPTY)) { skip = false; break; }
				}
				if (!skip)
				{
					counted = true;
					// cells of this block sample it and the first plane of its +1 neighbours (superset test)
					unsigned u = 0;
					for (int z = 0; z < 2; ++z) for (int y = 0; y < 2; ++y) for (int x = 0; x < 2; ++x)
					{
						const int cx = min(bx + x, nb - 1), cy = min(by + y, nb - 1), cz = min(bz + z, nb - 1);
						u |= d.blockInfo[((size_t)cz * nb + cy) * nb + cx];
					}
					take = (u & VXB_BI_NEG) && (u & VXB_BI_NONNEG);
				}
			}
			else if (level == 1)
			{
				// sign mix on the even lattice over the covered level-0 blocks + the far halo plane (superset);
				// blocks on the far grid edge sample the clamped coordinate n-1, which is off the lattice.
				unsigned u = 0; bool anyChild = false;
				const bool edge = (bx == nb - 1) || (by == nb - 1) || (bz == nb - 1);
				for (int z = 0; z < 3; ++z) for (int y = 0; y < 3; ++y) for (int x = 0; x < 3; ++x)
				{
					const int cx = min(2 * bx + x, nb0 - 1), cy = min(2 * by + y, nb0 - 1), cz = min(2 * bz + z, nb0 - 1);
					const size_t cb = ((size_t)cz * nb0 + cy) * nb0 + cx;
					const unsigned bi = d.blockInfo[cb];
					u |= (bi >> 3) & 3u;
					if (edge) u |= bi & 3u;
					if (x < 2 && y < 2 && z < 2 && d.consValid[cb]) anyChild = true;
				}
				take = (u == 3u) || anyChild;
			}
			else take = true;
		}
		const unsigned ballot = __ballot_sync(0xFFFFFFFFu, take);
		const unsigned cballot = __ballot_sync(0xFFFFFFFFu, counted);
		const unsigned lane = threadIdx.x & 31;
		unsigned off = 0;
		if (lane == 0)
		{
			if (ballot) off = atomicAdd(&d.counters->workCount[level], __popc(ballot));
			if (cballot) atomicAdd(&d.counters->nonSkippedLevel0, __popc(cballot));
		}
		off = __shfl_sync(0xFFFFFFFFu, off, 0);
		if (take) d.worklist[d.workBase[level] + off + __popc(ballot & ((1u << lane) - 1u))] = b;
	}
}

// ------------------------------------------------------------------------------------------------
// K3: polygonize
// ------------------------------------------------------------------------------------------------
// stage the 17^3 sample tile of block (bx,by,bz) at `level` into smem (TMA for level 0, clamped strided gather above)
__device__ __forceinline__ void vxb_stage_tile(signed char* tile, unsigned long long* mbar, unsigned& phase, const CUtensorMap* tmap,
	const VxbGrid& g, int n, int level, int bx, int by, int bz)
{
	const int tid = threadIdx.x;
	const int m = 1 << level, nb = n / 16 / m;
	if (level == 0)
	{
		if (tid == 0)
		{
			vxb_fence_proxy_async();
			vxb_mbar_expect_tx(mbar, VXB_TILE_BYTES);
			vxb_tma_load_3d(tile, tmap, bx * 16, by * 16, bz * 16, mbar);
		}
		vxb_mbar_wait(mbar, phase);
		phase ^= 1;
		// far grid edge: the +1 plane is outside the volume (TMA zero-fills); the reference clamps (:1037-1047)
		if (bx == nb - 1) { __syncthreads(); for (int i = tid; i < 17 * 17; i += VXB_THREADS) tile[i * VXB_TILE_PITCH + 16] = tile[i * VXB_TILE_PITCH + 15]; }
		if (by == nb - 1) { __syncthreads(); for (int i = tid; i < 17 * 17; i += VXB_THREADS) { const int z = i / 17, x = i % 17; tile[(z * 17 + 16) * VXB_TILE_PITCH + x] = tile[(z * 17 + 15) * VXB_TILE_PITCH + x]; } }
		if (bz == nb - 1) { __syncthreads(); for (int i = tid; i < 17 * 17; i += VXB_THREADS) { const int y = i / 17, x = i % 17; tile[(16 * 17 + y) * VXB_TILE_PITCH + x] = tile[(15 * 17 + y) * VXB_TILE_PITCH + x]; } }
	}
	else
	{
		const int lim = n - 1;
		for (int i = tid; i < 17 * 17 * 17; i += VXB_THREADS)
		{
			const int x = i % 17, y = (i / 17) % 17, z = i / 289;
			const int gx = min((bx * 16 + x) * m, lim), gy = min((by * 16 + y) * m, lim), gz = min((bz * 16 + z) * m, lim);
			tile[(z * 17 + y) * VXB_TILE_PITCH + x] = g.dist[((size_t)gz * n + gy) * n + gx];
		}
	}
}

struct __align__(128) VxbPolySmem
{
	signed char tile[VXB_TILE_BYTES + 96]; // 17 x 17 rows of 32 bytes; [z][y][x]
	unsigned int nt32[128];      // non-trivial cell bits, cell c -> word c>>5 bit c&31 (serial z,y,x order)
	unsigned int wpre[132];      // exclusive prefix of popc(nt32): compact index base of each word; [128] = total
	unsigned int recA[4096];     // per non-trivial cell (compact index): matId | matBlend<<8 | slotK<<16
	unsigned int recB[4096];     // vbase | newMask<<16
	unsigned short list[4096];   // compact index -> cell id
	unsigned int warpSums[8];
	unsigned int hist[16];
	unsigned int used[8];
	unsigned long long mbar;
	unsigned int item;
	unsigned int voff, ioff, tvoff, tioff;
	unsigned int removed;
	unsigned int hasChild;
	unsigned int pageReady;
	// transition scratch (one face at a time; cell = row*16 + col = thread id)
	unsigned char tslot[256][10];
	unsigned char tmat[256];
	unsigned short tnew[256];
	unsigned short tvbase[256];
	unsigned int tnt[8];
};

__device__ __forceinline__ unsigned vxb_rank(const VxbPolySmem& s, int c)
{
	return s.wpre[c >> 5] + __popc(s.nt32[c >> 5] & ((1u << (c & 31)) - 1u));
}

__device__ __forceinline__ void vxb_cell_samples(const VxbPolySmem& s, int c, signed char v[8])
{
	const int lx = c & 15, ly = (c >> 4) & 15, lz = c >> 8;
	const signed char* p = s.tile + (lz * 17 + ly) * VXB_TILE_PITCH + lx;
	v[0] = p[0]; v[1] = p[1]; v[2] = p[VXB_TILE_PITCH]; v[3] = p[VXB_TILE_PITCH + 1];
	p += 17 * VXB_TILE_PITCH;
	v[4] = p[0]; v[5] = p[1]; v[6] = p[VXB_TILE_PITCH]; v[7] = p[VXB_TILE_PITCH + 1];
}

// CalculateMaterialForCellCache :753-838 for level >= 1, children read through the page tables.
// The 8 children of a cell share one block of the child level (cell bases are even in child units), so the block
// lookup and the validity test happen once, and the common all-empty case exits after 4 loads.
__device__ __forceinline__ bool vxb_vote_cell(const VxbDev& d, int level, const int base[3], unsigned& id, unsigned& blend)
{
	const int cm = (1 << level) >> 1, cnb = d.n / 16 / cm;
	const int cx = base[0] / cm, cy = base[1] / cm, cz = base[2] / cm;            // child-level cell coordinates (even)
	const size_t bid = ((size_t)(cz >> 4) * cnb + (cy >> 4)) * cnb + (cx >> 4);
	const int lx = cx & 15, ly = cy & 15, lz = cz & 15;
	VxbVote v;
	vxb_vote_init(v);
	if (level == 1)
	{
		if (!d.consValid[bid]) return false;
		const unsigned short* rows = reinterpret_cast<const unsigned short*>(d.consPages + bid * 128); // 16 bits per (z,y) row
		const unsigned r00 = (rows[lz * 16 + ly] >> lx) & 3u, r01 = (rows[lz * 16 + ly + 1] >> lx) & 3u;
		const unsigned r10 = (rows[(lz + 1) * 16 + ly] >> lx) & 3u, r11 = (rows[(lz + 1) * 16 + ly + 1] >> lx) & 3u;
		const unsigned bits = r00 | (r01 << 2) | (r10 << 4) | (r11 << 6); // child q = x + 2y + 4z
		if (!bits) return false;
#pragma unroll
		for (int q = 0; q < 8; ++q)
		{
			unsigned cid = VXB_EMPTY_MATERIAL, cblend = 0;
			if ((bits >> q) & 1u)
			{
				const size_t gi = ((size_t)(base[2] + (q >> 2)) * d.n + (base[1] + ((q >> 1) & 1))) * d.n + (base[0] + (q & 1));
				cid = d.grid.mat[gi]; cblend = d.grid.blend[gi];
			}
			vxb_vote_add(v, cid, cblend);
		}
	}
	else
	{
		if (!d.cacheValid[level - 1][bid]) return false;
		const unsigned int* page = reinterpret_cast<const unsigned int*>(d.cachePages[level - 1] + bid * 4096); // 2 cells per word
		const int o = (lz * 256 + ly * 16 + lx) >> 1;
		const unsigned e0 = page[o], e1 = page[o + 8], e2 = page[o + 128], e3 = page[o + 136];
		if ((e0 & e1 & e2 & e3 & 0x00FF00FFu) == 0x00FF00FFu) return false; // all eight children EMPTY_MATERIAL
		const unsigned e[4] = { e0, e1, e2, e3 };
#pragma unroll
		for (int q = 0; q < 8; ++q)
		{
			const unsigned w = e[q >> 1] >> ((q & 1) * 16);
			vxb_vote_add(v, w & 0xFF, (w >> 8) & 0xFF);
		}
	}
	return vxb_vote_result(v, id, blend);
}

struct VxbDecision { bool isNew, quirkV0; unsigned ownerIdx; int ok; };

// new-vs-reuse decision for table vertex d of non-trivial cell c (compact index i) - :1610-1644
__device__ __forceinline__ VxbDecision vxb_decide(const VxbPolySmem& s, int c, int mask, const VxbVertexDesc& d, unsigned myMat)
{
	VxbDecision r; r.isNew = true; r.quirkV0 = false; r.ownerIdx = 0; r.ok = VXB_NO_SLOT;
	if (!d.atC7 && (d.dir & mask) == d.dir)
	{
		const int oc = c - (d.dir & 1) - ((d.dir >> 1) & 1) * 16 - ((d.dir >> 2) & 1) * 256;
		int ok = VXB_NO_SLOT; unsigned oi = 0, oa = 0;
		if ((s.nt32[oc >> 5] >> (oc & 31)) & 1u)
		{
			oi = vxb_rank(s, oc);
			oa = s.recA[oi];
			ok = (oa >> (16 + 4 * d.slot)) & 0xF;
		}
		if (ok != VXB_NO_SLOT)
		{
			if ((oa & 0xFF) == myMat) { r.isNew = false; r.ownerIdx = oi; r.ok = ok; }
		}
		else if (d.endpoint) r.quirkV0 = true;
	}
	return r;
}

__device__ __forceinline__ void vxb_store_vertex(VxbVertex* dst, const VxbVertex& v)
{
	const uint4* src = reinterpret_cast<const uint4*>(&v);
	uint4* out = reinterpret_cast<uint4*>(dst);
	out[0] = src[0]; out[1] = src[1]; out[2] = src[2];
}

// Generic (capacity-unbounded) kernel: processes the blocks the two emit tiers of vxb_emit.cuh rejected
// (> 4096... rather: > 12288 vertices in one block).  It redoes classification for its block (idempotent) and
// leaves the per-class / non-trivial statistics to vxb_classify_kernel.
__global__ void __launch_bounds__(VXB_THREADS) vxb_generic_kernel(const __grid_constant__ CUtensorMap tmap, const VxbDev d)
{
	extern __shared__ __align__(128) unsigned char smemRaw[];
	VxbPolySmem& s = *reinterpret_cast<VxbPolySmem*>(smemRaw);
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const VxbGrid g = d.grid;
	unsigned phase = 0;

	if (tid == 0) vxb_mbar_init(&s.mbar, 1);
	if (tid < 16) s.hist[tid] = 0;
	if (tid < 8) s.used[tid] = 0;
	unsigned statRemoved = 0; // thread 0 accumulates
	__syncthreads();

	const unsigned workCount = d.counters->genCount;
	for (;;)
	{
		if (tid == 0) s.item = atomicAdd(&d.counters->genCursor, 1u);
		__syncthreads();
		const unsigned item = s.item;
		if (item >= workCount) break;
		const unsigned packed = d.emitList[d.genList[item]];
		const int level = (int)(packed >> 28);
		const bool LEVEL0 = level == 0;
		const unsigned coordId = packed & 0x0FFFFFFFu;
		const int m = 1 << level, nb = d.n / 16 / m;
		const bool midLevel = level > 0 && level != d.lastLevel;
		const int bx = coordId % nb, by = (coordId / nb) % nb, bz = coordId / (nb * nb);

		// ---- step 0: stage the 17^3 sample tile ----
		vxb_stage_tile(s.tile, &s.mbar, phase, &tmap, g, d.n, level, bx, by, bz);
		if (tid == 0) { s.removed = 0; s.pageReady = 0; s.hasChild = 0; s.voff = 0; s.ioff = 0; }
		__syncthreads();

		// ---- step 1: case codes, non-trivial ballots (serial order: z slices, 256 cells each) ----
		unsigned myNt = 0; // bit z = my cell of slice z is non-trivial
#pragma unroll 4
		for (int z = 0; z < 16; ++z)
		{
			signed char v[8];
			const int c = z * 256 + tid;
			vxb_cell_samples(s, c, v);
			const unsigned code = vxb_case_code(v);
			const bool nt = (code != 0u && code != 255u);
			const unsigned bal = __ballot_sync(0xFFFFFFFFu, nt);
			if (lane == 0) s.nt32[z * 8 + warp] = bal;
			myNt |= (nt ? 1u : 0u) << z;
		}
		__syncthreads();
		unsigned ntc;
		{
			const unsigned cnt = (tid < 128) ? __popc(s.nt32[tid]) : 0u;
			const unsigned ex = vxb_block_scan(cnt, s.warpSums, ntc);
			if (tid < 128) s.wpre[tid] = ex;
			if (tid == 0) s.wpre[128] = ntc;
		}
		__syncthreads();

		unsigned nverts = 0, ntris = 0;
		if (ntc > 0)
		{
			// compact list in serial order
			for (int z = 0; z < 16; ++z)
				if ((myNt >> z) & 1u) { const int c = z * 256 + tid; s.list[vxb_rank(s, c)] = (unsigned short)c; }

			// pages: consistency bits (level 0) / material cache page (level >= 1)
			if (LEVEL0)
			{
				if (tid < 128)
				{
					unsigned* page = d.consPages + (size_t)coordId * 128;
					page[tid] = d.consValid[coordId] ? (page[tid] | s.nt32[tid]) : s.nt32[tid]; // bits are only ever set (:757)
				}
			}
			else if (!d.cacheValid[level][coordId])
			{
				unsigned int* page = reinterpret_cast<unsigned int*>(d.cachePages[level] + (size_t)coordId * 4096);
				for (int i = tid; i < 2048; i += VXB_THREADS) page[i] = 0x00FF00FFu; // {EMPTY_MATERIAL, 0} x 2  (:424)
			}
			__syncthreads();
			if (tid == 0)
			{
				if (LEVEL0) d.consValid[coordId] = 1; else d.cacheValid[level][coordId] = 1;
				s.pageReady = 1;
			}

			const unsigned per = (ntc + VXB_THREADS - 1) / VXB_THREADS;
			const unsigned i0 = min(tid * per, ntc), i1 = min(i0 + per, ntc);

			// ---- step 2: per non-trivial cell: material, table-vertex descriptors, owned slots ----
			for (unsigned i = i0; i < i1; ++i)
			{
				const int c = s.list[i];
				const int lx = c & 15, ly = (c >> 4) & 15, lz = c >> 8;
				const int base[3] = { (bx * 16 + lx) * m, (by * 16 + ly) * m, (bz * 16 + lz) * m };
				signed char v[8];
				vxb_cell_samples(s, c, v);
				const unsigned code = vxb_case_code(v);
				const unsigned cls = vxbRegularCellClass[code];
				unsigned matId = VXB_EMPTY_MATERIAL, matBlend = 0;
				if (LEVEL0)
				{
					const size_t gi = ((size_t)base[2] * d.n + base[1]) * d.n + base[0];
					matId = g.mat[gi]; matBlend = g.blend[gi];
				}
				else if (vxb_vote_cell(d, level, base, matId, matBlend))
					d.cachePages[level][(size_t)coordId * 4096 + c] = (unsigned short)(matId | (matBlend << 8));
				else { matId = VXB_EMPTY_MATERIAL; matBlend = 0; }
				unsigned slotK = 0xFFFFu;
				const int nv = vxbRegularCellData[cls * 16] >> 4;
				for (int k = 0; k < nv; ++k)
				{
					const VxbVertexDesc vd = vxb_regular_vertex_desc(vxbRegularVertexData[code * 12 + k], v);
					const int sl = vxb_regular_owned_slot(vd);
					if (sl >= 0) slotK = (slotK & ~(0xFu << (4 * sl))) | ((unsigned)k << (4 * sl));
				}
				s.recA[i] = matId | (matBlend << 8) | (slotK << 16);
			}
			__syncthreads();

			// ---- step 3: new-vs-reuse decisions -> counts ----
			unsigned myVerts = 0, myTris = 0;
			for (unsigned i = i0; i < i1; ++i)
			{
				const int c = s.list[i];
				signed char v[8];
				vxb_cell_samples(s, c, v);
				const unsigned code = vxb_case_code(v);
				const unsigned cls = vxbRegularCellClass[code];
				const unsigned geo = vxbRegularCellData[cls * 16];
				const unsigned rowStart = vxb_rank(s, c & ~15), sliceStart = s.wpre[(c >> 8) * 8];
				const int mask = (i > rowStart ? 1 : 0) | (rowStart > sliceStart ? 2 : 0) | (sliceStart > 0 ? 4 : 0);
				const unsigned myMat = s.recA[i] & 0xFF;
				unsigned newMask = 0;
				for (int k = 0; k < (int)(geo >> 4); ++k)
				{
					const VxbVertexDesc vd = vxb_regular_vertex_desc(vxbRegularVertexData[code * 12 + k], v);
					if (vxb_decide(s, c, mask, vd, myMat).isNew) newMask |= 1u << k;
				}
				s.recB[i] = newMask << 16;
				myVerts += __popc(newMask);
				myTris += geo & 0xF;
			}
			unsigned packedTotal;
			const unsigned packedBase = vxb_block_scan(myVerts | (myTris << 16), s.warpSums, packedTotal);
			nverts = packedTotal & 0xFFFF; ntris = packedTotal >> 16;
			{
				unsigned vb = packedBase & 0xFFFF;
				for (unsigned i = i0; i < i1; ++i) { const unsigned nm = s.recB[i] >> 16; s.recB[i] = vb | (nm << 16); vb += __popc(nm); }
			}
			if (tid == 0)
			{
				s.voff = atomicAdd(&d.counters->vertices, nverts);
				s.ioff = atomicAdd(&d.counters->indices, ntris * 3);
			}
			__syncthreads();
			const unsigned voff = s.voff, ioff = s.ioff;
			const bool fits = (unsigned long long)voff + nverts <= d.vcap && (unsigned long long)ioff + ntris * 3ull <= d.icap;

			if (fits)
			{
				// ---- step 4: emit the new vertices ----
				for (unsigned i = i0; i < i1; ++i)
				{
					const int c = s.list[i];
					const int lx = c & 15, ly = (c >> 4) & 15, lz = c >> 8;
					const int local[3] = { lx, ly, lz };
					const int base[3] = { (bx * 16 + lx) * m, (by * 16 + ly) * m, (bz * 16 + lz) * m };
					signed char v[8];
					vxb_cell_samples(s, c, v);
					const unsigned code = vxb_case_code(v);
					const unsigned cls = vxbRegularCellClass[code];
					const unsigned ra = s.recA[i], rb = s.recB[i];
					const unsigned matId = ra & 0xFF, matBlend = (ra >> 8) & 0xFF, newMask = rb >> 16;
					if (!newMask) continue;
					const unsigned rowStart = vxb_rank(s, c & ~15), sliceStart = s.wpre[(c >> 8) * 8];
					const int mask = (i > rowStart ? 1 : 0) | (rowStart > sliceStart ? 2 : 0) | (sliceStart > 0 ? 4 : 0);
					unsigned vid = voff + (rb & 0xFFFF);
					atomicOr(&s.used[matId >> 5], 1u << (matId & 31));
					const int nv = vxbRegularCellData[cls * 16] >> 4;
					for (int k = 0; k < nv; ++k)
					{
						if (!((newMask >> k) & 1u)) continue;
						const VxbVertexDesc vd = vxb_regular_vertex_desc(vxbRegularVertexData[code * 12 + k], v);
						VxbRawVertex rv;
						if (vd.endpoint)
						{
							const bool quirk = vxb_decide(s, c, mask, vd, matId).quirkV0;
							vxb_corner_vertex(g, level, base, local, quirk ? vd.v0 : ((vd.t == 0) ? vd.v1 : vd.v0), matId, matBlend, rv);
						}
						else vxb_edge_vertex(g, level, base, local, vd, matId, matBlend, rv);
						vxb_regular_secondary(level, rv);
						VxbVertex ov;
						vxb_finish_vertex(rv, *d.lut, ov);
						vxb_store_vertex(d.verts + vid, ov);
						++vid;
					}
				}
				__syncthreads(); // vertices of this block are visible to the whole CTA

				// ---- step 5: triangles + degenerate filter (:1300-1321); removed ones are marked, compacted below ----
				unsigned tb = packedBase >> 16, myRemoved = 0;
				for (unsigned i = i0; i < i1; ++i)
				{
					const int c = s.list[i];
					signed char v[8];
					vxb_cell_samples(s, c, v);
					const unsigned code = vxb_case_code(v);
					const unsigned cls = vxbRegularCellClass[code];
					const unsigned geo = vxbRegularCellData[cls * 16];
					const unsigned ra = s.recA[i], rb = s.recB[i];
					const unsigned rowStart = vxb_rank(s, c & ~15), sliceStart = s.wpre[(c >> 8) * 8];
					const int mask = (i > rowStart ? 1 : 0) | (rowStart > sliceStart ? 2 : 0) | (sliceStart > 0 ? 4 : 0);
					unsigned vids[12];
					unsigned nextNew = rb & 0xFFFF;
					for (int k = 0; k < (int)(geo >> 4); ++k)
					{
						if ((rb >> (16 + k)) & 1u) { vids[k] = nextNew++; continue; }
						const VxbVertexDesc vd = vxb_regular_vertex_desc(vxbRegularVertexData[code * 12 + k], v);
						const VxbDecision dec = vxb_decide(s, c, mask, vd, ra & 0xFF);
						const unsigned ob = s.recB[dec.ownerIdx];
						vids[k] = (ob & 0xFFFF) + __popc((ob >> 16) & ((1u << dec.ok) - 1u));
					}
					for (unsigned tr = 0; tr < (geo & 0xF); ++tr, ++tb)
					{
						const unsigned a = vids[vxbRegularCellData[cls * 16 + 1 + tr * 3]];
						const unsigned b = vids[vxbRegularCellData[cls * 16 + 2 + tr * 3]];
						const unsigned cc = vids[vxbRegularCellData[cls * 16 + 3 + tr * 3]];
						float pa[3], pb[3], pc[3];
						{
							const float* fa = d.verts[voff + a].pos; const float* fb = d.verts[voff + b].pos; const float* fc = d.verts[voff + cc].pos;
							// back to grid axes, x256 (exact: positions are multiples of 1/256)
							pa[0] = fa[0] * 256.f; pa[1] = fa[2] * 256.f; pa[2] = fa[1] * 256.f;
							pb[0] = fb[0] * 256.f; pb[1] = fb[2] * 256.f; pb[2] = fb[1] * 256.f;
							pc[0] = fc[0] * 256.f; pc[1] = fc[2] * 256.f; pc[2] = fc[1] * 256.f;
						}
						unsigned* out = d.idx + ioff + tb * 3;
						if (vxb_triangle_kept(pa, pb, pc)) { out[0] = a; out[1] = b; out[2] = cc; }
						else { out[0] = 0xFFFFFFFFu; out[1] = 0xFFFFFFFFu; out[2] = 0xFFFFFFFFu; ++myRemoved; }
					}
				}
				if (myRemoved) atomicAdd(&s.removed, myRemoved);
				__syncthreads();
				const unsigned removed = s.removed;
				if (removed)
				{
					// order-preserving in-place compaction of this block's triangle list
					unsigned written = 0;
					for (unsigned t0 = 0; t0 < ntris; t0 += VXB_THREADS)
					{
						const unsigned t = t0 + tid;
						unsigned a = 0xFFFFFFFFu, b = 0, cc = 0;
						if (t < ntris) { const unsigned* in = d.idx + ioff + t * 3; a = in[0]; b = in[1]; cc = in[2]; }
						const bool keep = (t < ntris) && a != 0xFFFFFFFFu;
						unsigned chunkTotal;
						const unsigned pos = vxb_block_scan(keep ? 1u : 0u, s.warpSums, chunkTotal); // syncs: all reads of this chunk are done
						if (keep) { unsigned* out = d.idx + ioff + (written + pos) * 3; out[0] = a; out[1] = b; out[2] = cc; }
						written += chunkTotal;
						__syncthreads();
					}
				}
				if (tid == 0) { statRemoved += removed; }
			}
		}

		// ---- step 6: transition cells (:1754-2131) + their material-cache side effect (:1859) ----
		unsigned tvCount[6] = { 0, 0, 0, 0, 0, 0 }, tiCount[6] = { 0, 0, 0, 0, 0, 0 }, tvOff[6] = { 0, 0, 0, 0, 0, 0 }, tiOff[6] = { 0, 0, 0, 0, 0, 0 };
		if (midLevel)
		{
			if (tid < 8)
			{
				const int cnb = nb * 2;
				const size_t cb = ((size_t)(bz * 2 + (tid >> 2)) * cnb + (by * 2 + ((tid >> 1) & 1))) * cnb + (bx * 2 + (tid & 1));
				const bool valid = (level == 1) ? d.consValid[cb] : d.cacheValid[level - 1][cb];
				if (valid) atomicOr(&s.hasChild, 1u);
			}
			__syncthreads();
			const bool hasChild = s.hasChild != 0;
			const bool outputTrans = d.transitions && nverts > 0 && ((unsigned long long)s.voff + nverts <= d.vcap);
			if (hasChild || outputTrans)
			{
				const int row = tid >> 4, col = tid & 15;
				for (int face = 0; face < 6; ++face)
				{
					int axis, ua, va;
					vxb_face_axes(face, axis, ua, va);
					const int bc = (axis == 0) ? bx : (axis == 1 ? by : bz);
					if (face < 3 ? (bc == 0) : (bc == nb - 1)) continue; // neighbour block outside the grid (:1829-1835)
					int local[3];
					local[axis] = (face >= 3) ? 15 : 0; local[ua] = col; local[va] = row;
					const int c = local[2] * 256 + local[1] * 16 + local[0];
					const int base[3] = { (bx * 16 + local[0]) * m, (by * 16 + local[1]) * m, (bz * 16 + local[2]) * m };

					unsigned matId = VXB_EMPTY_MATERIAL, matBlend = 0;
					bool voted = false;
					if (hasChild) voted = vxb_vote_cell(d, level, base, matId, matBlend);
					if (!voted) { matId = VXB_EMPTY_MATERIAL; matBlend = 0; }
					if (!s.pageReady) // block-uniform
					{
						if (__syncthreads_or(voted ? 1 : 0))
						{
							if (!d.cacheValid[level][coordId])
							{
								unsigned int* page = reinterpret_cast<unsigned int*>(d.cachePages[level] + (size_t)coordId * 4096);
								for (int i = tid; i < 2048; i += VXB_THREADS) page[i] = 0x00FF00FFu;
							}
							__syncthreads();
							if (tid == 0) { d.cacheValid[level][coordId] = 1; s.pageReady = 1; }
							__syncthreads();
						}
					}
					if (voted) d.cachePages[level][(size_t)coordId * 4096 + c] = (unsigned short)(matId | (matBlend << 8));
					if (!outputTrans) continue;

					signed char v[13];
#pragma unroll
					for (int i = 0; i < 9; ++i)
					{
						int p[3];
						vxb_transition_sample_pos(face, level, base, i, p);
						v[i] = (signed char)vxb_dist(g, p[0], p[1], p[2]);
					}
					v[9] = v[0]; v[10] = v[2]; v[11] = v[6]; v[12] = v[8];
					const unsigned code = vxb_transition_case_code(v);
					const bool nt = (code != 0u && code != 511u);
					const unsigned bal = __ballot_sync(0xFFFFFFFFu, nt);
					if (lane == 0) s.tnt[warp] = bal;
					const unsigned cls = nt ? vxbTransitionCellClass[code] : 0u;
					const unsigned char* cd = &vxbTransitionCellData[(cls & 0x7F) * 40];
					const int nv = nt ? (cd[0] >> 4) : 0, ntri = nt ? (cd[0] & 0xF) : 0;
					s.tmat[tid] = (unsigned char)matId;
#pragma unroll
					for (int q = 0; q < 10; ++q) s.tslot[tid][q] = VXB_NO_SLOT;
					const int mask = ((row > 0) ? 2 : 0) | ((bal >> (lane & 16)) & ((1u << col) - 1u) ? 1 : 0);
					for (int k = 0; k < nv; ++k)
					{
						const VxbTransVertexDesc td = vxb_transition_vertex_desc(vxbTransitionVertexData[code * 12 + k], v, vxbTransitionCornerData);
						if ((td.dir & mask) != td.dir && td.dir == 8) s.tslot[tid][td.slot] = (unsigned char)k; // stored only when no reuse was attempted (:2097)
					}
					__syncthreads();
					unsigned newMask = 0;
					unsigned ownerOf[12];
					for (int k = 0; k < nv; ++k)
					{
						const VxbTransVertexDesc td = vxb_transition_vertex_desc(vxbTransitionVertexData[code * 12 + k], v, vxbTransitionCornerData);
						bool isNew = true;
						ownerOf[k] = 0;
						if ((td.dir & mask) == td.dir)
						{
							const int oc = (row - ((td.dir >> 1) & 1)) * 16 + (col - (td.dir & 1));
							const bool ont = (s.tnt[oc >> 5] >> (oc & 31)) & 1u;
							const int ok = ont ? s.tslot[oc][td.slot] : VXB_NO_SLOT;
							if (ok != VXB_NO_SLOT && s.tmat[oc] == matId) { isNew = false; ownerOf[k] = (unsigned)oc | ((unsigned)ok << 8); }
						}
						if (isNew) newMask |= 1u << k;
					}
					s.tnew[tid] = (unsigned short)newMask;
					unsigned packedTotal;
					const unsigned packedBase = vxb_block_scan(__popc(newMask) | ((unsigned)ntri << 16), s.warpSums, packedTotal);
					const unsigned fv = packedTotal & 0xFFFF, ft = packedTotal >> 16;
					s.tvbase[tid] = (unsigned short)(packedBase & 0xFFFF);
					if (tid == 0)
					{
						s.tvoff = fv ? atomicAdd(&d.counters->transVertices, fv) : 0u;
						s.tioff = ft ? atomicAdd(&d.counters->transIndices, ft * 3) : 0u;
					}
					__syncthreads();
					const unsigned tvoff = s.tvoff, tioff = s.tioff;
					const bool tfits = (unsigned long long)tvoff + fv <= d.tvcap && (unsigned long long)tioff + ft * 3ull <= d.ticap;
					tvCount[face] = fv; tiCount[face] = ft * 3; tvOff[face] = tvoff; tiOff[face] = tioff;
					if (nt && tfits)
					{
						unsigned vids[12];
						unsigned nextNew = packedBase & 0xFFFF;
						for (int k = 0; k < nv; ++k)
						{
							if ((newMask >> k) & 1u)
							{
								const VxbTransVertexDesc td = vxb_transition_vertex_desc(vxbTransitionVertexData[code * 12 + k], v, vxbTransitionCornerData);
								VxbRawVertex rv;
								vxb_transition_vertex(g, face, level, base, local, td, matId, matBlend, rv);
								VxbVertex ov;
								vxb_finish_vertex(rv, *d.lut, ov);
								vxb_store_vertex(d.tverts + tvoff + nextNew, ov);
								vids[k] = nextNew++;
							}
							else
							{
								const unsigned oc = ownerOf[k] & 0xFF, ok = ownerOf[k] >> 8;
								vids[k] = s.tvbase[oc] + __popc((unsigned)s.tnew[oc] & ((1u << ok) - 1u));
							}
						}
						atomicOr(&s.used[matId >> 5], 1u << (matId & 31));
						const bool flip = (((cls >> 7) & 1u) ^ (unsigned)(face & 1)) != 0;
						unsigned* out = d.tidx + tioff + (packedBase >> 16) * 3;
						for (int tr = 0; tr < ntri; ++tr)
						{
							const unsigned a = vids[cd[1 + tr * 3]], b = vids[cd[2 + tr * 3]], cc = vids[cd[3 + tr * 3]];
							out[tr * 3] = a; out[tr * 3 + 1] = flip ? cc : b; out[tr * 3 + 2] = flip ? b : cc;
						}
					}
					__syncthreads(); // scratch is reused by the next face
				}
			}
		}

		// ---- step 7: directory record (PushBlocksToResult: only blocks with >= 1 vertex :1278) ----
		if (tid == 0 && nverts > 0)
		{
			const unsigned slot = atomicAdd(&d.counters->records, 1u);
			if (slot < d.rcap)
			{
				vxb_block_record r;
				r.level = level; r.coord_id = coordId; r.id = d.idBase[level] + coordId;
				r.vertex_count = nverts; r.index_count = (ntris - s.removed) * 3;
				r.vertex_offset = s.voff; r.index_offset = s.ioff;
				for (int f = 0; f < 6; ++f)
				{
					// internal face order (z-,y-,x-,z+,y+,x+) is already the output enum order (YNeg,ZNeg,XNeg,YPos,ZPos,XPos)
					r.trans_vertex_count[f] = tvCount[f]; r.trans_index_count[f] = tiCount[f];
					r.trans_vertex_offset[f] = tvOff[f]; r.trans_index_offset[f] = tiOff[f];
				}
				r.reserved = 0;
				d.records[slot] = r;
			}
		}
		__syncthreads(); // tile / scratch are free for the next item
	}

	// ---- statistics ----
	__syncthreads();
	if (tid < 8 && s.used[tid]) atomicOr(&d.counters->usedMaterials[tid], s.used[tid]);
	if (tid == 0 && statRemoved) atomicAdd(&d.counters->degenerate, statRemoved);
}
