// Consumer-side packaging: the result stays in HBM and is handed to a renderer as indirect draws.
// The arenas already are merged vertex / index buffers (every block a contiguous range, indices block-local), so a
// draw of a block is {index_count, first_index, base_vertex}.  What the renderer contract of the reference adds
// (doc_source/Rendering.md:18-58) is per frame: WHICH blocks to draw (a cut through the LOD octree: finer near the
// camera), for each drawn block of level > 0 which of its six faces border a finer-level neighbour - there the block's
// transition mesh is drawn and the vertex shader swaps in the secondary position of the vertices whose mask
// (SecondaryPosition.w) is covered by that face set (`blockAdj`, Rendering.md:44-52).
#pragma once

struct VxbLodArgs
{
	float camera[3];      // output (Y-up) coordinates, like the vertices
	float baseDistance;   // a node of level l is fine enough when the camera is farther than baseDistance * 2^l from its centre
	unsigned int* counts; // [0] regular draws, [1] transition draws
	vxb_draw_command* regular; vxb_draw_info* regularInfo;
	vxb_draw_command* transition; vxb_draw_info* transitionInfo;
	unsigned int capacity; // entries of each list
};

// does the octree node (level, grid block coordinates) stop the descent?  (the top level always does)
__device__ __forceinline__ bool vxb_lod_stops(const VxbDev& d, const VxbLodArgs& a, int level, int bx, int by, int bz)
{
	if (level == d.lastLevel) return true;
	const float m = (float)(16 << level);
	// node centre in output axes: (x, z, y) (:1289-1291)
	const float cx = ((float)bx + 0.5f) * m, cy = ((float)bz + 0.5f) * m, cz = ((float)by + 0.5f) * m;
	const float dx = cx - a.camera[0], dy = cy - a.camera[1], dz = cz - a.camera[2];
	const float dist2 = (dx * dx + dy * dy) + dz * dz;
	const float lim = a.baseDistance * (float)(1 << level);
	return dist2 >= lim * lim;
}

// is the node part of the cut: it stops and every ancestor descends
__device__ __forceinline__ bool vxb_lod_selected(const VxbDev& d, const VxbLodArgs& a, int level, int bx, int by, int bz)
{
	if (level > 0 && !vxb_lod_stops(d, a, level, bx, by, bz)) return false; // level 0 cannot descend further
	for (int l = level + 1; l <= d.lastLevel; ++l)
		if (vxb_lod_stops(d, a, l, bx >> (l - level), by >> (l - level), bz >> (l - level))) return false;
	return true;
}

// is the region of node (level, b) drawn with FINER blocks: the node descends and so does every ancestor
__device__ __forceinline__ bool vxb_lod_finer(const VxbDev& d, const VxbLodArgs& a, int level, int bx, int by, int bz)
{
	for (int l = level; l <= d.lastLevel; ++l)
		if (vxb_lod_stops(d, a, l, bx >> (l - level), by >> (l - level), bz >> (l - level))) return false;
	return level > 0;
}

// one thread per emitted block (directory slot)
__global__ void __launch_bounds__(256) vxb_lod_select_kernel(const VxbDev d, const VxbLodArgs a, const unsigned int blockCount)
{
	const unsigned slot = blockIdx.x * blockDim.x + threadIdx.x;
	if (slot >= blockCount) return;
	const vxb_block_record r = d.records[slot];
	const int level = (int)r.level, nb = (d.n >> 4) >> level;
	const int bx = r.coord_id % nb, by = (r.coord_id / nb) % nb, bz = r.coord_id / (nb * nb);
	if (!vxb_lod_selected(d, a, level, bx, by, bz)) return;
	// faces in the output enum order = internal order z-, y-, x-, z+, y+, x+ (vxb_finish_kernel)
	unsigned adj = 0;
	if (level > 0)
	{
		const int dxyz[6][3] = { { 0, 0, -1 }, { 0, -1, 0 }, { -1, 0, 0 }, { 0, 0, 1 }, { 0, 1, 0 }, { 1, 0, 0 } };
		for (int f = 0; f < 6; ++f)
		{
			const int nx = bx + dxyz[f][0], ny = by + dxyz[f][1], nz = bz + dxyz[f][2];
			if (nx < 0 || ny < 0 || nz < 0 || nx >= nb || ny >= nb || nz >= nb) continue;
			if (vxb_lod_finer(d, a, level, nx, ny, nz)) adj |= 1u << f;
		}
	}
	const unsigned k = atomicAdd(&a.counts[0], 1u);
	if (k < a.capacity)
	{
		vxb_draw_command c;
		c.index_count = r.index_count; c.instance_count = 1; c.first_index = r.index_offset; c.base_vertex = (int)r.vertex_offset; c.first_instance = k;
		a.regular[k] = c;
		vxb_draw_info inf;
		inf.block_id = r.id; inf.level = r.level; inf.block_adj = adj; inf.face = 0xFFFFFFFFu;
		a.regularInfo[k] = inf;
	}
	for (int f = 0; f < 6; ++f)
	{
		if (!((adj >> f) & 1u) || !r.trans_index_count[f]) continue;
		const unsigned t = atomicAdd(&a.counts[1], 1u);
		if (t >= a.capacity) continue;
		vxb_draw_command c;
		c.index_count = r.trans_index_count[f]; c.instance_count = 1; c.first_index = r.trans_index_offset[f]; c.base_vertex = (int)r.trans_vertex_offset[f]; c.first_instance = t;
		a.transition[t] = c;
		vxb_draw_info inf;
		inf.block_id = r.id; inf.level = r.level; inf.block_adj = adj; inf.face = (unsigned)f;
		a.transitionInfo[t] = inf;
	}
}
