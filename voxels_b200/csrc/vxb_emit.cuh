// Fast path of the polygonizer (included by vxb200.cu after vxb_kernels.cuh).
//
//   vxb_classify_kernel   per LOD level (levels depend on each other ONLY through this step):
//                         tile -> case codes -> non-trivial bits; consistency page (level 0) or majority votes of
//                         the non-trivial cells and of every transition-face cell into the level's material page
//                         (CalculateMaterialForCellCache :753-838, called from :1568 and :1859); statistics;
//                         appends blocks that have non-trivial cells to the emit list together with their bit mask.
//   vxb_emit_kernel<C,V>  ONE launch for the blocks of ALL levels (largest blocks first): ordered compaction, reuse
//                         decisions, block scan, thread-per-new-vertex emission, triangles + degenerate filter,
//                         all six transition faces at once.  Works out of ~40 KB of shared memory for blocks with
//                         <= C non-trivial cells and <= V vertices; anything larger is forwarded to the next tier
//                         (C,V = 4096,12288) and finally to the generic kernel of vxb_kernels.cuh.
//
// Bit-exactness notes are in vxb_kernels.cuh / vxb_cell.h; the formulation is identical, only the schedule differs.
#pragma once

struct __align__(128) VxbClassifySmem
{
	signed char tile[VXB_TILE_BYTES + 96];
	unsigned int rowSign[17 * 17 + 3]; // bit x = sample (x, y, z) of the tile is negative
	unsigned int nt32[128];
	unsigned int wpre[132];
	unsigned short list[4096];
	unsigned int warpSums[8];
	unsigned int hist[16];
	unsigned long long mbar;
	unsigned int item, hasChild, pageReady, emitIdx;
};

// Non-trivial bits of all 4096 cells from the tile, 16 cells per thread-step instead of one:
// a cell is non-trivial iff its 8 corner signs are neither all 0 nor all 1 (Cell::CalcCaseCode :741-750, :1560).
__device__ __forceinline__ void vxb_classify_bits(const signed char* tile, unsigned int* rowSign, unsigned int* nt32)
{
	const int tid = threadIdx.x;
	for (int r = tid; r < 17 * 17; r += VXB_THREADS)
	{
		const unsigned int* w = reinterpret_cast<const unsigned int*>(tile + r * VXB_TILE_PITCH);
		unsigned m = 0;
#pragma unroll
		for (int q = 0; q < 4; ++q) m |= ((((w[q] >> 7) & 0x01010101u) * 0x01020408u) >> 24 & 0xFu) << (4 * q);
		m |= ((w[4] >> 7) & 1u) << 16;
		rowSign[r] = m;
	}
	__syncthreads();
	{
		const int z = tid >> 4, y = tid & 15;
		const unsigned a = rowSign[z * 17 + y], b = rowSign[z * 17 + y + 1], c = rowSign[(z + 1) * 17 + y], e = rowSign[(z + 1) * 17 + y + 1];
		const unsigned any = a | b | c | e, all = a & b & c & e;
		reinterpret_cast<unsigned short*>(nt32)[tid] = (unsigned short)((any | (any >> 1)) & ~(all & (all >> 1)) & 0xFFFFu);
	}
}

__device__ __forceinline__ void vxb_tile_samples(const signed char* tile, int c, signed char v[8])
{
	const int lx = c & 15, ly = (c >> 4) & 15, lz = c >> 8;
	const signed char* p = tile + (lz * 17 + ly) * VXB_TILE_PITCH + lx;
	v[0] = p[0]; v[1] = p[1]; v[2] = p[VXB_TILE_PITCH]; v[3] = p[VXB_TILE_PITCH + 1];
	p += 17 * VXB_TILE_PITCH;
	v[4] = p[0]; v[5] = p[1]; v[6] = p[VXB_TILE_PITCH]; v[7] = p[VXB_TILE_PITCH + 1];
}

__device__ __forceinline__ void vxb_init_cache_page(const VxbDev& d, int level, unsigned coordId)
{
	unsigned int* page = reinterpret_cast<unsigned int*>(d.cachePages[level] + (size_t)coordId * 4096);
	for (int i = threadIdx.x; i < 2048; i += VXB_THREADS) page[i] = 0x00FF00FFu; // {EMPTY_MATERIAL, 0} x 2  (:424)
}

__global__ void __launch_bounds__(VXB_THREADS, 4) vxb_classify_kernel(const __grid_constant__ CUtensorMap tmap, const VxbDev d, const int level)
{
	extern __shared__ __align__(128) unsigned char smemRaw[];
	VxbClassifySmem& s = *reinterpret_cast<VxbClassifySmem*>(smemRaw);
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int m = 1 << level, nb = d.n / 16 / m;
	const bool midLevel = level > 0 && level != d.lastLevel;
	unsigned phase = 0;
	if (tid == 0) vxb_mbar_init(&s.mbar, 1);
	if (tid < 16) s.hist[tid] = 0;
	unsigned statNonTrivial = 0;
	__syncthreads();

	const unsigned workCount = d.counters->workCount[level];
	for (;;)
	{
		if (tid == 0) { s.item = atomicAdd(&d.counters->workCursor[level], 1u); s.hasChild = 0; s.pageReady = 0; }
		__syncthreads();
		const unsigned item = s.item;
		if (item >= workCount) break;
		const unsigned coordId = d.worklist[d.workBase[level] + item];
		const int bx = coordId % nb, by = (coordId / nb) % nb, bz = coordId / (nb * nb);
		vxb_stage_tile(s.tile, &s.mbar, phase, &tmap, d.grid, d.n, level, bx, by, bz);
		__syncthreads();

		vxb_classify_bits(s.tile, s.rowSign, s.nt32);
		__syncthreads();
		unsigned ntc;
		{
			const unsigned cnt = (tid < 128) ? __popc(s.nt32[tid]) : 0u;
			const unsigned ex = vxb_block_scan(cnt, s.warpSums, ntc);
			if (tid < 128) s.wpre[tid] = ex;
		}
		__syncthreads();

		if (ntc > 0)
		{
			if (level == 0)
			{
				if (tid < 128)
				{
					unsigned* page = d.consPages + (size_t)coordId * 128;
					page[tid] = d.consValid[coordId] ? (page[tid] | s.nt32[tid]) : s.nt32[tid]; // bits are only ever set (:757)
				}
			}
			else if (!d.cacheValid[level][coordId]) vxb_init_cache_page(d, level, coordId);
			{
				// ordered compact list: thread = cell row (z, y), 16 bits each
				unsigned bits = reinterpret_cast<const unsigned short*>(s.nt32)[tid];
				unsigned pos = s.wpre[tid >> 1] + ((tid & 1) ? __popc(s.nt32[tid >> 1] & 0xFFFFu) : 0u);
				while (bits) { const int x = __ffs(bits) - 1; bits &= bits - 1; s.list[pos++] = (unsigned short)(tid * 16 + x); }
			}
			__syncthreads();
			if (tid == 0)
			{
				if (level == 0) d.consValid[coordId] = 1; else d.cacheValid[level][coordId] = 1;
				s.pageReady = 1;
				const unsigned idx = d.workBase[level] + atomicAdd(&d.counters->emitCount[level], 1u);
				d.emitList[idx] = ((unsigned)level << 28) | coordId;
				s.emitIdx = idx;
				statNonTrivial += ntc;
			}
			for (unsigned i = tid; i < ntc; i += VXB_THREADS)
			{
				const int c = s.list[i];
				signed char v[8];
				vxb_tile_samples(s.tile, c, v);
				atomicAdd(&s.hist[vxbGRegularCellClass[vxb_case_code(v)]], 1u); // PerCaseCellsCount (:1574)
				if (level > 0)
				{
					// material of every non-trivial cell (:1568): majority vote of its 8 children, stored in the level's page
					const int base[3] = { (bx * 16 + (c & 15)) * m, (by * 16 + ((c >> 4) & 15)) * m, (bz * 16 + (c >> 8)) * m };
					unsigned matId, matBlend;
					if (vxb_vote_cell(d, level, base, matId, matBlend))
						d.cachePages[level][(size_t)coordId * 4096 + c] = (unsigned short)(matId | (matBlend << 8));
				}
			}
			__syncthreads();
			if (tid < 128) d.ntScratch[(size_t)s.emitIdx * 128 + tid] = s.nt32[tid];
		}

		// every cell of every transition face votes too, trivial ones included (:1859) - observable at the next level
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
			if (s.hasChild)
			{
				const int row = tid >> 4, col = tid & 15;
				unsigned votes[6]; // matId | blend<<8 | ok<<16
				bool any = false;
#pragma unroll
				for (int face = 0; face < 6; ++face)
				{
					votes[face] = 0;
					int axis, ua, va;
					vxb_face_axes(face, axis, ua, va);
					const int bc = (axis == 0) ? bx : (axis == 1 ? by : bz);
					if (face < 3 ? (bc == 0) : (bc == nb - 1)) continue; // neighbour block outside the grid (:1829-1835)
					int local[3];
					local[axis] = (face >= 3) ? 15 : 0; local[ua] = col; local[va] = row;
					const int base[3] = { (bx * 16 + local[0]) * m, (by * 16 + local[1]) * m, (bz * 16 + local[2]) * m };
					unsigned matId, matBlend;
					if (vxb_vote_cell(d, level, base, matId, matBlend)) { votes[face] = matId | (matBlend << 8) | (1u << 16); any = true; }
				}
				if (!s.pageReady) // block-uniform
				{
					if (__syncthreads_or(any ? 1 : 0))
					{
						if (!d.cacheValid[level][coordId]) vxb_init_cache_page(d, level, coordId);
						__syncthreads();
						if (tid == 0) { d.cacheValid[level][coordId] = 1; s.pageReady = 1; }
					}
				}
				if (any)
				{
#pragma unroll
					for (int face = 0; face < 6; ++face)
					{
						if (!(votes[face] >> 16)) continue;
						int axis, ua, va;
						vxb_face_axes(face, axis, ua, va);
						int local[3];
						local[axis] = (face >= 3) ? 15 : 0; local[ua] = col; local[va] = row;
						d.cachePages[level][(size_t)coordId * 4096 + local[2] * 256 + local[1] * 16 + local[0]] = (unsigned short)(votes[face] & 0xFFFF);
					}
				}
			}
		}
		__syncthreads();
	}

	__syncthreads();
	if (tid < 16 && s.hist[tid]) atomicAdd(&d.counters->perCase[tid], s.hist[tid]);
	if (tid == 0 && statNonTrivial) atomicAdd(&d.counters->nonTrivial, statNonTrivial);
}

// ------------------------------------------------------------------------------------------------
// emit
// ------------------------------------------------------------------------------------------------
template <int CAP_C, int CAP_V>
struct __align__(128) VxbEmitSmem
{
	signed char tile[VXB_TILE_BYTES + 96];
	unsigned int nt32[128];
	unsigned int wpre[132];
	unsigned char tabClass[256];
	unsigned char tabCell[256];
	unsigned short tabVert[3072];
	unsigned int warpSums[8];
	unsigned int used[8];
	unsigned long long mbar;
	unsigned int item, voff, ioff, removed;
	unsigned int tvoff[6], tioff[6];
	unsigned int levelEnd[VXB_MAX_LEVELS + 1];
	union
	{
		struct
		{
			unsigned int recA[CAP_C];     // matId | matBlend<<8 | slotK<<16
			unsigned int recB[CAP_C];     // newMask | quirkMask<<12
			unsigned short list[CAP_C];   // compact index -> cell id
			unsigned short cz[CAP_C];     // case code | zero mask << 8
			unsigned short vbase[CAP_C];  // exclusive scan of new-vertex counts
			unsigned short tbase[CAP_C];  // exclusive scan of triangle counts
			unsigned short vlist[CAP_V];  // new vertex -> compact cell index << 4 | table vertex
		} r;
		struct // all six transition faces at once: cell = face*256 + row*16 + col
		{
			unsigned long long slots[1536]; // 10 owned-slot nibbles per cell
			unsigned short newMask[1536];
			unsigned short vbase[1536];
			unsigned char mat[1536];
			unsigned int nt[48];
		} t;
	} u;
};

template <int CAP_C, int CAP_V>
__device__ __forceinline__ unsigned vxb_emit_rank(const VxbEmitSmem<CAP_C, CAP_V>& s, int c)
{
	return s.wpre[c >> 5] + __popc(s.nt32[c >> 5] & ((1u << (c & 31)) - 1u));
}

// new-vs-reuse decision (:1610-1644) against the owner's record
template <int CAP_C, int CAP_V>
__device__ __forceinline__ VxbDecision vxb_emit_decide(const VxbEmitSmem<CAP_C, CAP_V>& s, int c, int mask, const VxbVertexDesc& d, unsigned myMat)
{
	VxbDecision r; r.isNew = true; r.quirkV0 = false; r.ownerIdx = 0; r.ok = VXB_NO_SLOT;
	if (!d.atC7 && (d.dir & mask) == d.dir)
	{
		const int oc = c - (d.dir & 1) - ((d.dir >> 1) & 1) * 16 - ((d.dir >> 2) & 1) * 256;
		int ok = VXB_NO_SLOT; unsigned oi = 0, oa = 0;
		if ((s.nt32[oc >> 5] >> (oc & 31)) & 1u)
		{
			oi = vxb_emit_rank(s, oc);
			oa = s.u.r.recA[oi];
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

template <int CAP_C, int CAP_V>
__device__ __forceinline__ int vxb_emit_mask(const VxbEmitSmem<CAP_C, CAP_V>& s, int c, unsigned i)
{
	const unsigned rowStart = vxb_emit_rank(s, c & ~15), sliceStart = s.wpre[(c >> 8) * 8];
	return (i > rowStart ? 1 : 0) | (rowStart > sliceStart ? 2 : 0) | (sliceStart > 0 ? 4 : 0);
}

// TIER 0: items come from the emit list (all levels, top level first), rejects go to bigList.
// TIER 1: items come from bigList, rejects go to genList (generic kernel).
template <int CAP_C, int CAP_V, int TIER>
__global__ void __launch_bounds__(VXB_THREADS, (CAP_C <= 1024 ? 4 : 2)) vxb_emit_kernel(const __grid_constant__ CUtensorMap tmap, const VxbDev d)
{
	typedef VxbEmitSmem<CAP_C, CAP_V> Smem;
	extern __shared__ __align__(128) unsigned char smemRaw[];
	Smem& s = *reinterpret_cast<Smem*>(smemRaw);
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const VxbGrid g = d.grid;
	unsigned phase = 0;

	if (tid == 0)
	{
		vxb_mbar_init(&s.mbar, 1);
		unsigned acc = 0; // work id -> level: the top level first (its blocks are the largest)
		for (int l = d.levels - 1; l >= 0; --l) { acc += d.counters->emitCount[l]; s.levelEnd[d.levels - 1 - l] = acc; }
	}
	if (tid < 8) s.used[tid] = 0;
	for (int i = tid; i < 256; i += VXB_THREADS) { s.tabClass[i] = vxbGRegularCellClass[i]; s.tabCell[i] = vxbGRegularCellData[i]; }
	for (int i = tid; i < 3072; i += VXB_THREADS) s.tabVert[i] = vxbGRegularVertexData[i];
	unsigned statRemoved = 0;
	__syncthreads();
	const unsigned workCount = (TIER == 0) ? s.levelEnd[d.levels - 1] : d.counters->bigCount;

	for (;;)
	{
		if (tid == 0) { s.item = atomicAdd(TIER == 0 ? &d.counters->emitCursor : &d.counters->bigCursor, 1u); s.removed = 0; }
		__syncthreads();
		const unsigned item = s.item;
		if (item >= workCount) break;
		unsigned emitIdx;
		if (TIER == 0)
		{
			int q = 0;
			while (item >= s.levelEnd[q]) ++q;
			const int lv = d.levels - 1 - q;
			emitIdx = d.workBase[lv] + (item - (q ? s.levelEnd[q - 1] : 0u));
		}
		else emitIdx = d.bigList[item];
		const unsigned packed = d.emitList[emitIdx];
		const int level = (int)(packed >> 28);
		const unsigned coordId = packed & 0x0FFFFFFFu;
		const int m = 1 << level, nb = d.n / 16 / m;
		const int bx = coordId % nb, by = (coordId / nb) % nb, bz = coordId / (nb * nb);
		const bool midLevel = level > 0 && level != d.lastLevel;

		if (tid < 128) s.nt32[tid] = d.ntScratch[(size_t)emitIdx * 128 + tid];
		vxb_stage_tile(s.tile, &s.mbar, phase, &tmap, g, d.n, level, bx, by, bz);
		__syncthreads();
		unsigned ntc;
		{
			const unsigned cnt = (tid < 128) ? __popc(s.nt32[tid]) : 0u;
			const unsigned ex = vxb_block_scan(cnt, s.warpSums, ntc);
			if (tid < 128) s.wpre[tid] = ex;
		}
		__syncthreads();
		bool reject = ntc > (unsigned)CAP_C;
		unsigned nverts = 0, ntris = 0;

		if (!reject)
		{
			// ordered compact list of the non-trivial cells: thread = cell row (z, y), 16 bits each
			{
				unsigned bits = reinterpret_cast<const unsigned short*>(s.nt32)[tid];
				unsigned pos = s.wpre[tid >> 1] + ((tid & 1) ? __popc(s.nt32[tid >> 1] & 0xFFFFu) : 0u);
				while (bits) { const int x = __ffs(bits) - 1; bits &= bits - 1; s.u.r.list[pos++] = (unsigned short)(tid * 16 + x); }
			}
			__syncthreads();

			// ---- pass A: material, descriptors, owned slots ----
			for (unsigned i = tid; i < ntc; i += VXB_THREADS)
			{
				const int c = s.u.r.list[i];
				signed char v[8];
				vxb_tile_samples(s.tile, c, v);
				const unsigned code = vxb_case_code(v);
				const unsigned cls = s.tabClass[code];
				unsigned matId, matBlend;
				if (level == 0)
				{
					const size_t gi = ((size_t)((bz * 16 + (c >> 8))) * d.n + (by * 16 + ((c >> 4) & 15))) * d.n + (bx * 16 + (c & 15));
					matId = g.mat[gi]; matBlend = g.blend[gi];
				}
				else
				{
					const unsigned e = d.cachePages[level][(size_t)coordId * 4096 + c]; // written by vxb_classify_kernel
					matId = e & 0xFF; matBlend = e >> 8;
				}
				unsigned slotK = 0xFFFFu;
				const unsigned zm = vxb_zero_mask(v);
				const int nv = s.tabCell[cls * 16] >> 4;
				for (int k = 0; k < nv; ++k)
				{
					const VxbVertexDesc vd = vxb_regular_vertex_desc_lite(s.tabVert[code * 12 + k], zm);
					const int sl = vxb_regular_owned_slot(vd);
					if (sl >= 0) slotK = (slotK & ~(0xFu << (4 * sl))) | ((unsigned)k << (4 * sl));
				}
				s.u.r.cz[i] = (unsigned short)(code | (zm << 8));
				s.u.r.recA[i] = matId | (matBlend << 8) | (slotK << 16);
			}
			__syncthreads();

			// ---- pass B: new-vs-reuse decisions ----
			for (unsigned i = tid; i < ntc; i += VXB_THREADS)
			{
				const int c = s.u.r.list[i];
				const unsigned code = s.u.r.cz[i] & 0xFF, zm = s.u.r.cz[i] >> 8;
				const unsigned geo = s.tabCell[s.tabClass[code] * 16];
				const int mask = vxb_emit_mask(s, c, i);
				const unsigned myMat = s.u.r.recA[i] & 0xFF;
				unsigned newMask = 0, quirkMask = 0;
				for (int k = 0; k < (int)(geo >> 4); ++k)
				{
					const VxbVertexDesc vd = vxb_regular_vertex_desc_lite(s.tabVert[code * 12 + k], zm);
					const VxbDecision dec = vxb_emit_decide(s, c, mask, vd, myMat);
					if (dec.isNew) newMask |= 1u << k;
					if (dec.quirkV0) quirkMask |= 1u << k;
				}
				s.u.r.recB[i] = newMask | (quirkMask << 12);
				s.u.r.vbase[i] = (unsigned short)__popc(newMask);
				s.u.r.tbase[i] = (unsigned short)(geo & 0xF);
			}
			__syncthreads();

			// ---- exclusive scans in serial cell order (contiguous chunk per thread) ----
			{
				const unsigned per = (ntc + VXB_THREADS - 1) / VXB_THREADS;
				const unsigned i0 = min(tid * per, ntc), i1 = min(i0 + per, ntc);
				unsigned sv = 0, st = 0;
				for (unsigned i = i0; i < i1; ++i) { sv += s.u.r.vbase[i]; st += s.u.r.tbase[i]; }
				unsigned total;
				const unsigned base = vxb_block_scan(sv | (st << 16), s.warpSums, total);
				nverts = total & 0xFFFF; ntris = total >> 16;
				unsigned bv = base & 0xFFFF, bt = base >> 16;
				if (nverts <= (unsigned)CAP_V || CAP_V >= 49152)
					for (unsigned i = i0; i < i1; ++i)
					{
						const unsigned cv = s.u.r.vbase[i], ct = s.u.r.tbase[i];
						s.u.r.vbase[i] = (unsigned short)bv; s.u.r.tbase[i] = (unsigned short)bt;
						bv += cv; bt += ct;
					}
			}
			reject = nverts > (unsigned)CAP_V;
		}

		if (reject)
		{
			if (tid == 0)
			{
				if (TIER == 0) d.bigList[atomicAdd(&d.counters->bigCount, 1u)] = emitIdx;
				else d.genList[atomicAdd(&d.counters->genCount, 1u)] = emitIdx;
			}
			__syncthreads();
			continue;
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
			// new-vertex list: vertex j -> (compact cell, table vertex)
			for (unsigned i = tid; i < ntc; i += VXB_THREADS)
			{
				unsigned nm = s.u.r.recB[i] & 0xFFFu, j = s.u.r.vbase[i];
				while (nm) { const int k = __ffs(nm) - 1; nm &= nm - 1; s.u.r.vlist[j++] = (unsigned short)((i << 4) | k); }
			}
			__syncthreads();

			// ---- pass C: one thread per new vertex ----
			for (unsigned j = tid; j < nverts; j += VXB_THREADS)
			{
				const unsigned e = s.u.r.vlist[j];
				const unsigned i = e >> 4; const int k = e & 15;
				const int c = s.u.r.list[i];
				const int local[3] = { c & 15, (c >> 4) & 15, c >> 8 };
				const int base[3] = { (bx * 16 + local[0]) * m, (by * 16 + local[1]) * m, (bz * 16 + local[2]) * m };
				const unsigned code = s.u.r.cz[i] & 0xFF, zm = s.u.r.cz[i] >> 8;
				const unsigned ra = s.u.r.recA[i];
				const unsigned matId = ra & 0xFF, matBlend = (ra >> 8) & 0xFF;
				VxbVertexDesc vd = vxb_regular_vertex_desc_lite(s.tabVert[code * 12 + k], zm);
				if (!vd.endpoint && level == 0)
				{
					const signed char* p = s.tile + (local[2] * 17 + local[1]) * VXB_TILE_PITCH + local[0];
					const int a = p[(vd.v0 & 1) + ((vd.v0 >> 1) & 1) * VXB_TILE_PITCH + (vd.v0 >> 2) * 17 * VXB_TILE_PITCH];
					const int b = p[(vd.v1 & 1) + ((vd.v1 >> 1) & 1) * VXB_TILE_PITCH + (vd.v1 >> 2) * 17 * VXB_TILE_PITCH];
					vd.t = vxb_fixed_t(a, b); // :1591 (coarser levels recompute t after the LOD descent)
				}
				VxbRawVertex rv;
				if (vd.endpoint)
				{
					const bool quirk = (s.u.r.recB[i] >> (12 + k)) & 1u;
					vxb_corner_vertex(g, level, base, local, quirk ? vd.v0 : ((vd.t == 0) ? vd.v1 : vd.v0), matId, matBlend, rv);
				}
				else vxb_edge_vertex(g, level, base, local, vd, matId, matBlend, rv);
				vxb_regular_secondary(level, rv);
				VxbVertex ov;
				vxb_finish_vertex(rv, *d.lut, ov);
				vxb_store_vertex(d.verts + voff + j, ov);
				atomicOr(&s.used[matId >> 5], 1u << (matId & 31));
			}
			__syncthreads(); // the block's vertices are visible to the whole CTA

			// ---- pass D: triangles + degenerate filter (:1300-1321) ----
			unsigned myRemoved = 0;
			for (unsigned i = tid; i < ntc; i += VXB_THREADS)
			{
				const int c = s.u.r.list[i];
				const unsigned code = s.u.r.cz[i] & 0xFF, zm = s.u.r.cz[i] >> 8;
				const unsigned cls = s.tabClass[code];
				const unsigned geo = s.tabCell[cls * 16];
				const unsigned ra = s.u.r.recA[i], rb = s.u.r.recB[i];
				const int mask = vxb_emit_mask(s, c, i);
				unsigned vids[12];
				unsigned nextNew = s.u.r.vbase[i];
				for (int k = 0; k < (int)(geo >> 4); ++k)
				{
					if ((rb >> k) & 1u) { vids[k] = nextNew++; continue; }
					const VxbVertexDesc vd = vxb_regular_vertex_desc_lite(s.tabVert[code * 12 + k], zm);
					const VxbDecision dec = vxb_emit_decide(s, c, mask, vd, ra & 0xFF);
					vids[k] = s.u.r.vbase[dec.ownerIdx] + __popc(s.u.r.recB[dec.ownerIdx] & 0xFFFu & ((1u << dec.ok) - 1u));
				}
				unsigned* out = d.idx + ioff + (unsigned)s.u.r.tbase[i] * 3;
				for (unsigned tr = 0; tr < (geo & 0xF); ++tr, out += 3)
				{
					const unsigned a = vids[s.tabCell[cls * 16 + 1 + tr * 3]];
					const unsigned b = vids[s.tabCell[cls * 16 + 2 + tr * 3]];
					const unsigned cc = vids[s.tabCell[cls * 16 + 3 + tr * 3]];
					const float* fa = d.verts[voff + a].pos; const float* fb = d.verts[voff + b].pos; const float* fc = d.verts[voff + cc].pos;
					// back to grid axes, x256 (exact: positions are multiples of 1/256)
					const float pa[3] = { fa[0] * 256.f, fa[2] * 256.f, fa[1] * 256.f };
					const float pb[3] = { fb[0] * 256.f, fb[2] * 256.f, fb[1] * 256.f };
					const float pc[3] = { fc[0] * 256.f, fc[2] * 256.f, fc[1] * 256.f };
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
					const unsigned pos = vxb_block_scan(keep ? 1u : 0u, s.warpSums, chunkTotal);
					if (keep) { unsigned* out = d.idx + ioff + (written + pos) * 3; out[0] = a; out[1] = b; out[2] = cc; }
					written += chunkTotal;
					__syncthreads();
				}
				if (tid == 0) statRemoved += removed;
			}
		}

		// ---- transition cells, all six faces at once (:1754-2131); materials come from the level's page ----
		unsigned tvCount[6] = { 0, 0, 0, 0, 0, 0 }, tiCount[6] = { 0, 0, 0, 0, 0, 0 };
		if (midLevel && d.transitions && fits)
		{
			__syncthreads(); // the regular-cell arrays are dead: their storage becomes the transition scratch
			const int row = tid >> 4, col = tid & 15;
			unsigned codes[6], newMasks[6];
			unsigned activeFaces = 0;
#pragma unroll
			for (int face = 0; face < 6; ++face)
			{
				int axis, ua, va;
				vxb_face_axes(face, axis, ua, va);
				const int bc = (axis == 0) ? bx : (axis == 1 ? by : bz);
				codes[face] = 0; newMasks[face] = 0;
				const bool active = !(face < 3 ? (bc == 0) : (bc == nb - 1)); // neighbour block inside the grid (:1829-1835)
				if (active) activeFaces |= 1u << face;
				bool nt = false;
				if (active)
				{
					int local[3];
					local[axis] = (face >= 3) ? 15 : 0; local[ua] = col; local[va] = row;
					const int base[3] = { (bx * 16 + local[0]) * m, (by * 16 + local[1]) * m, (bz * 16 + local[2]) * m };
					signed char v[9];
#pragma unroll
					for (int q = 0; q < 9; ++q)
					{
						int p[3];
						vxb_transition_sample_pos(face, level, base, q, p);
						v[q] = (signed char)vxb_dist(g, p[0], p[1], p[2]);
					}
					codes[face] = vxb_transition_case_code(v);
					nt = codes[face] != 0u && codes[face] != 511u;
				}
				const unsigned bal = __ballot_sync(0xFFFFFFFFu, nt);
				if (lane == 0) s.u.t.nt[face * 8 + warp] = bal;
				if (!nt) codes[face] = 0;
			}
			// owned slots + material per non-trivial transition cell
#pragma unroll
			for (int face = 0; face < 6; ++face)
			{
				const int ci = face * 256 + tid;
				unsigned long long slots = ~0ull;
				unsigned matId = VXB_EMPTY_MATERIAL;
				if (codes[face])
				{
					int axis, ua, va;
					vxb_face_axes(face, axis, ua, va);
					int local[3];
					local[axis] = (face >= 3) ? 15 : 0; local[ua] = col; local[va] = row;
					const int c = local[2] * 256 + local[1] * 16 + local[0];
					matId = d.cachePages[level][(size_t)coordId * 4096 + c] & 0xFF;
					const int base[3] = { (bx * 16 + local[0]) * m, (by * 16 + local[1]) * m, (bz * 16 + local[2]) * m };
					signed char v[13];
					for (int q = 0; q < 9; ++q) { int p[3]; vxb_transition_sample_pos(face, level, base, q, p); v[q] = (signed char)vxb_dist(g, p[0], p[1], p[2]); }
					v[9] = v[0]; v[10] = v[2]; v[11] = v[6]; v[12] = v[8];
					const unsigned code = codes[face];
					const int nv = vxbGTransitionCellData[(vxbGTransitionCellClass[code] & 0x7F) * 40] >> 4;
					for (int k = 0; k < nv; ++k)
					{
						const VxbTransVertexDesc td = vxb_transition_vertex_desc(vxbGTransitionVertexData[code * 12 + k], v, vxbGTransitionCornerData);
						if (td.dir == 8) slots = (slots & ~(0xFull << (4 * td.slot))) | ((unsigned long long)k << (4 * td.slot)); // stored only when no reuse was attempted (:2097)
					}
				}
				s.u.t.slots[ci] = slots;
				s.u.t.mat[ci] = (unsigned char)matId;
			}
			__syncthreads();
			// decisions
			unsigned counts[6];
#pragma unroll
			for (int face = 0; face < 6; ++face)
			{
				counts[face] = 0;
				if (!codes[face]) { s.u.t.newMask[face * 256 + tid] = 0; continue; }
				int axis, ua, va;
				vxb_face_axes(face, axis, ua, va);
				int local[3];
				local[axis] = (face >= 3) ? 15 : 0; local[ua] = col; local[va] = row;
				const int base[3] = { (bx * 16 + local[0]) * m, (by * 16 + local[1]) * m, (bz * 16 + local[2]) * m };
				signed char v[13];
				for (int q = 0; q < 9; ++q) { int p[3]; vxb_transition_sample_pos(face, level, base, q, p); v[q] = (signed char)vxb_dist(g, p[0], p[1], p[2]); }
				v[9] = v[0]; v[10] = v[2]; v[11] = v[6]; v[12] = v[8];
				const unsigned code = codes[face];
				const unsigned cls = vxbGTransitionCellClass[code];
				const unsigned geo = vxbGTransitionCellData[(cls & 0x7F) * 40];
				const unsigned rowBits = (s.u.t.nt[face * 8 + warp] >> (lane & 16)) & 0xFFFFu;
				const int mask = ((row > 0) ? 2 : 0) | ((rowBits & ((1u << col) - 1u)) ? 1 : 0);
				const unsigned myMat = s.u.t.mat[face * 256 + tid];
				unsigned newMask = 0;
				for (int k = 0; k < (int)(geo >> 4); ++k)
				{
					const VxbTransVertexDesc td = vxb_transition_vertex_desc(vxbGTransitionVertexData[code * 12 + k], v, vxbGTransitionCornerData);
					bool isNew = true;
					if ((td.dir & mask) == td.dir)
					{
						const int oc = face * 256 + (row - ((td.dir >> 1) & 1)) * 16 + (col - (td.dir & 1));
						const int ok = (int)((s.u.t.slots[oc] >> (4 * td.slot)) & 0xF); // trivial cells hold all-NO_SLOT
						if (ok != VXB_NO_SLOT && s.u.t.mat[oc] == myMat) isNew = false;
					}
					if (isNew) newMask |= 1u << k;
				}
				newMasks[face] = newMask;
				s.u.t.newMask[face * 256 + tid] = (unsigned short)newMask;
				counts[face] = __popc(newMask) | ((geo & 0xF) << 16);
			}
			// per-face ordered scans (thread order = row-major cell order)
			unsigned bases[6];
#pragma unroll
			for (int face = 0; face < 6; ++face)
			{
				bases[face] = 0;
				if (!((activeFaces >> face) & 1u)) continue; // block-uniform
				unsigned total;
				bases[face] = vxb_block_scan(counts[face], s.warpSums, total);
				tvCount[face] = total & 0xFFFF; tiCount[face] = (total >> 16) * 3;
				s.u.t.vbase[face * 256 + tid] = (unsigned short)(bases[face] & 0xFFFF);
			}
			if (tid < 6)
			{
				s.tvoff[tid] = tvCount[tid] ? atomicAdd(&d.counters->transVertices, tvCount[tid]) : 0u;
				s.tioff[tid] = tiCount[tid] ? atomicAdd(&d.counters->transIndices, tiCount[tid]) : 0u;
			}
			__syncthreads();
#pragma unroll
			for (int face = 0; face < 6; ++face)
			{
				if (!codes[face]) continue;
				const unsigned tvoff = s.tvoff[face], tioff = s.tioff[face];
				if ((unsigned long long)tvoff + tvCount[face] > d.tvcap || (unsigned long long)tioff + tiCount[face] > d.ticap) continue;
				int axis, ua, va;
				vxb_face_axes(face, axis, ua, va);
				int local[3];
				local[axis] = (face >= 3) ? 15 : 0; local[ua] = col; local[va] = row;
				const int c = local[2] * 256 + local[1] * 16 + local[0];
				const int base[3] = { (bx * 16 + local[0]) * m, (by * 16 + local[1]) * m, (bz * 16 + local[2]) * m };
				signed char v[13];
				for (int q = 0; q < 9; ++q) { int p[3]; vxb_transition_sample_pos(face, level, base, q, p); v[q] = (signed char)vxb_dist(g, p[0], p[1], p[2]); }
				v[9] = v[0]; v[10] = v[2]; v[11] = v[6]; v[12] = v[8];
				const unsigned code = codes[face];
				const unsigned cls = vxbGTransitionCellClass[code];
				const unsigned char* cd = &vxbGTransitionCellData[(cls & 0x7F) * 40];
				const int nv = cd[0] >> 4, ntri = cd[0] & 0xF;
				const unsigned e = d.cachePages[level][(size_t)coordId * 4096 + c];
				const unsigned matId = e & 0xFF, matBlend = e >> 8;
				const unsigned rowBits = (s.u.t.nt[face * 8 + warp] >> (lane & 16)) & 0xFFFFu;
				const int mask = ((row > 0) ? 2 : 0) | ((rowBits & ((1u << col) - 1u)) ? 1 : 0);
				unsigned vids[12];
				unsigned nextNew = bases[face] & 0xFFFF;
				for (int k = 0; k < nv; ++k)
				{
					const VxbTransVertexDesc td = vxb_transition_vertex_desc(vxbGTransitionVertexData[code * 12 + k], v, vxbGTransitionCornerData);
					if ((newMasks[face] >> k) & 1u)
					{
						VxbRawVertex rv;
						vxb_transition_vertex(g, face, level, base, local, td, matId, matBlend, rv);
						VxbVertex ov;
						vxb_finish_vertex(rv, *d.lut, ov);
						vxb_store_vertex(d.tverts + tvoff + nextNew, ov);
						vids[k] = nextNew++;
					}
					else
					{
						const int oc = face * 256 + (row - ((td.dir >> 1) & 1)) * 16 + (col - (td.dir & 1));
						const unsigned ok = (unsigned)((s.u.t.slots[oc] >> (4 * td.slot)) & 0xF);
						vids[k] = s.u.t.vbase[oc] + __popc((unsigned)s.u.t.newMask[oc] & ((1u << ok) - 1u));
					}
				}
				atomicOr(&s.used[matId >> 5], 1u << (matId & 31));
				const bool flip = (((cls >> 7) & 1u) ^ (unsigned)(face & 1)) != 0;
				unsigned* out = d.tidx + tioff + (bases[face] >> 16) * 3;
				for (int tr = 0; tr < ntri; ++tr)
				{
					const unsigned a = vids[cd[1 + tr * 3]], b = vids[cd[2 + tr * 3]], cc = vids[cd[3 + tr * 3]];
					out[tr * 3] = a; out[tr * 3 + 1] = flip ? cc : b; out[tr * 3 + 2] = flip ? b : cc;
				}
			}
		}

		// ---- directory record (PushBlocksToResult: blocks with >= 1 vertex, :1278) ----
		if (tid == 0)
		{
			const unsigned slot = atomicAdd(&d.counters->records, 1u);
			if (slot < d.rcap)
			{
				vxb_block_record r;
				r.level = level; r.coord_id = coordId; r.id = d.idBase[level] + coordId;
				r.vertex_count = nverts; r.index_count = (ntris - s.removed) * 3;
				r.vertex_offset = voff; r.index_offset = ioff;
				for (int f = 0; f < 6; ++f)
				{
					r.trans_vertex_count[f] = tvCount[f]; r.trans_index_count[f] = tiCount[f];
					r.trans_vertex_offset[f] = tvCount[f] ? s.tvoff[f] : 0u; r.trans_index_offset[f] = tiCount[f] ? s.tioff[f] : 0u;
				}
				r.reserved = 0;
				d.records[slot] = r;
			}
		}
		__syncthreads();
	}

	__syncthreads();
	if (tid < 8 && s.used[tid]) atomicOr(&d.counters->usedMaterials[tid], s.used[tid]);
	if (tid == 0 && statRemoved) atomicAdd(&d.counters->degenerate, statRemoved);
}
