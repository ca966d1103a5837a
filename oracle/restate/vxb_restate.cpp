// TEST INFRASTRUCTURE - CPU restatement of the reference's polygonization path
// (/root/reference/src/TransVoxelImpl.cpp) in the PARALLEL formulation the CUDA kernels use:
// per 16^3 block  A) classify every cell independently,  B) decide new-vs-reused per table vertex
// from local facts only,  C) exclusive scan in serial cell order + emit  (SURVEY.md Appendix C).
// It shares the per-cell arithmetic (voxels_b200/csrc/vxb_cell.h) with the kernels and replaces
// warps/ballots/scans by plain loops, so it is the debuggable middle rung between the real
// reference (oracle/_ref, which pins it - see tests/test_restate_vs_reference.py) and the GPU.
// Nothing in voxels_b200/ may link or call this file.
//
// Reference anchors: Execute :468-538, GenerateBlockListForLevel :385-466, PolygonizeBlock :1529-1750,
// CalculateMaterialForCellCache :753-838, GenerateTransitionCells :1754-2131, PushBlocksToResult :1266-1428,
// AreBlockAndNeighborsEmpty :1511-1527, VoxelGrid::CompressBlock (BF_Empty rule) VoxelGrid.cpp:610-672.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <vector>
#include <algorithm>

#define VXB_TABLE_QUAL static
#include "vxb_tables_data.h"
#include "../../voxels_b200/csrc/vxb_cell.h"

namespace
{

struct OutBlock
{
	uint32_t id;
	uint32_t coordId;
	float mn[3], mx[3];
	std::vector<VxbVertex> verts;
	std::vector<uint32_t> idx;
	std::vector<VxbVertex> tverts[6];
	std::vector<uint32_t> tidx[6];
};

struct CacheEntry { uint8_t id, blend; };

struct Run
{
	VxbGrid grid;
	VxbMaterialLut lut;
	int n, levels;
	std::vector<uint8_t> emptyFlags;                       // per L0 block: BF_Empty
	std::map<uint32_t, std::vector<uint32_t>> cons;        // L0 coordId -> 128 x u32 consistency bits
	std::vector<std::map<uint32_t, std::vector<CacheEntry>>> cache; // [level-1]: coordId -> 4096 entries
	std::vector<std::vector<OutBlock>> out;                // per level, z,y,x order
	uint32_t stats[20];
	uint32_t nextId;
};

// VoxelGrid::CompressBlock's isEmpty / compressionEffective outcome for one block (VoxelGrid.cpp:610-672)
bool blockIsEmpty(const VxbGrid& g, int bx, int by, int bz)
{
	signed char data[4096];
	for (int z = 0; z < 16; ++z) for (int y = 0; y < 16; ++y)
		memcpy(data + z * 256 + y * 16, g.dist + ((size_t)(bz * 16 + z) * g.n + (by * 16 + y)) * g.n + bx * 16, 16);
	bool isEmpty = true;
	unsigned counter = 0, runEnds = 0;
	const int initial = data[0];
	signed char last = data[0];
	for (unsigned i = 0; i < 4096; ++i)
	{
		const signed char cur = data[i];
		if (last == cur && counter < 0xFF) { ++counter; continue; }
		++runEnds;
		counter = 1;
		last = cur;
		if (initial * last <= 0) isEmpty = false;
		if (1 + 2 * runEnds > 4096) return false; // compression ineffective -> stored raw, never "empty"
	}
	return isEmpty;
}

inline uint32_t coordId(int nb, int bx, int by, int bz) { return (uint32_t)((bz * nb + by) * nb + bx); }

// ---- material of a cell (CalculateMaterialForCellCache :753-838) -----------------------------------
// Level 0 is handled by the caller.  Returns false when no child carried a material.
bool voteCell(Run& r, int level, const int base[3], unsigned& id, unsigned& blend)
{
	const int cm = (1 << level) >> 1;          // child multiplier
	const int cnb = r.n / 16 / cm;             // child level block count per axis
	VxbVote v;
	vxb_vote_init(v);
	for (int z = 0; z < 2; ++z) for (int y = 0; y < 2; ++y) for (int x = 0; x < 2; ++x)
	{
		const int c[3] = { base[0] + x * cm, base[1] + y * cm, base[2] + z * cm };
		const int ext = 16 * cm;
		const uint32_t bid = coordId(cnb, c[0] / ext, c[1] / ext, c[2] / ext);
		const unsigned lid = (unsigned)(((c[2] % ext) / cm) * 256 + ((c[1] % ext) / cm) * 16 + ((c[0] % ext) / cm));
		unsigned cid = VXB_EMPTY_MATERIAL, cblend = 0;
		if (level == 1)
		{
			auto it = r.cons.find(bid);
			if (it != r.cons.end() && ((it->second[lid >> 5] >> (lid & 31)) & 1))
			{
				const size_t gi = vxb_index(r.grid, c[0], c[1], c[2]);
				cid = r.grid.mat[gi]; cblend = r.grid.blend[gi];
			}
		}
		else
		{
			auto& lv = r.cache[level - 2];
			auto it = lv.find(bid);
			if (it != lv.end()) { cid = it->second[lid].id; cblend = it->second[lid].blend; }
		}
		vxb_vote_add(v, cid, cblend);
	}
	return vxb_vote_result(v, id, blend);
}

void cacheStore(Run& r, int level, uint32_t bid, unsigned lid, unsigned id, unsigned blend)
{
	auto& lv = r.cache[level - 1];
	auto it = lv.find(bid);
	if (it == lv.end())
		it = lv.emplace(bid, std::vector<CacheEntry>(4096, CacheEntry{ VXB_EMPTY_MATERIAL, 0 })).first;
	it->second[lid].id = (uint8_t)id;
	it->second[lid].blend = (uint8_t)blend;
}

struct CellRec
{
	signed char v[8];
	uint8_t code, nontrivial, matId, matBlend;
	uint16_t slotK;    // 4 nibbles, VXB_NO_SLOT = none
	uint16_t newMask;
	uint32_t vbase;
};

// ---- one block: regular cells ------------------------------------------------------------------
void processBlock(Run& r, int level, int bx, int by, int bz, bool withTransitions, OutBlock& ob)
{
	const int m = 1 << level, nb = r.n / 16 / m;
	const uint32_t bid = coordId(nb, bx, by, bz);
	const VxbGrid& g = r.grid;
	std::vector<CellRec> cells(4096);

	// A: classify
	for (int c = 0; c < 4096; ++c)
	{
		CellRec& cr = cells[c];
		const int lx = c & 15, ly = (c >> 4) & 15, lz = c >> 8;
		const int base[3] = { (bx * 16 + lx) * m, (by * 16 + ly) * m, (bz * 16 + lz) * m };
		for (int i = 0; i < 8; ++i)
			cr.v[i] = (signed char)vxb_dist(g, base[0] + ((i & 1) ? m : 0), base[1] + ((i & 2) ? m : 0), base[2] + ((i & 4) ? m : 0));
		cr.code = (uint8_t)vxb_case_code(cr.v);
		cr.nontrivial = (cr.code != 0 && cr.code != 255);
		cr.slotK = 0xFFFF; cr.newMask = 0; cr.vbase = 0; cr.matId = VXB_EMPTY_MATERIAL; cr.matBlend = 0;
		if (!cr.nontrivial) { ++r.stats[1]; continue; }
		++r.stats[2];
		const unsigned cls = vxbRegularCellClass[cr.code];
		++r.stats[4 + cls];
		if (level == 0)
		{
			auto it = r.cons.find(bid);
			if (it == r.cons.end()) it = r.cons.emplace(bid, std::vector<uint32_t>(128, 0u)).first;
			it->second[c >> 5] |= 1u << (c & 31);
			const size_t gi = vxb_index(g, base[0], base[1], base[2]);
			cr.matId = g.mat[gi]; cr.matBlend = g.blend[gi];
		}
		else
		{
			unsigned id, blend;
			if (voteCell(r, level, base, id, blend)) { cr.matId = (uint8_t)id; cr.matBlend = (uint8_t)blend; cacheStore(r, level, bid, c, id, blend); }
			// else: the reference leaves cell.Material uninitialised (UB); we define it as {EMPTY, 0}
		}
		const int nverts = vxbRegularCellData[cls * 16] >> 4;
		for (int k = 0; k < nverts; ++k)
		{
			const VxbVertexDesc d = vxb_regular_vertex_desc(vxbRegularVertexData[cr.code * 12 + k], cr.v);
			const int s = vxb_regular_owned_slot(d);
			if (s >= 0) cr.slotK = (uint16_t)((cr.slotK & ~(0xF << (4 * s))) | (k << (4 * s)));
		}
	}

	// B: new-vs-reuse decisions from local facts; C: scan + emit
	struct Ref { int owner; int ok; bool quirkV0; };
	std::vector<VxbRawVertex> raw;
	std::vector<uint32_t> rawIdx;
	bool rowSeen[256] = { false }, sliceSeen[16] = { false };
	for (int c = 0; c < 4096; ++c) if (cells[c].nontrivial) { rowSeen[c >> 4] = true; sliceSeen[c >> 8] = true; }
	uint32_t vcount = 0;
	for (int c = 0; c < 4096; ++c)
	{
		CellRec& cr = cells[c];
		if (!cr.nontrivial) continue;
		const int lx = c & 15, ly = (c >> 4) & 15, lz = c >> 8;
		int mask = 0;
		for (int x = 0; x < lx; ++x) if (cells[c - lx + x].nontrivial) mask |= 1;
		for (int y = 0; y < ly; ++y) if (rowSeen[lz * 16 + y]) mask |= 2;
		for (int z = 0; z < lz; ++z) if (sliceSeen[z]) mask |= 4;

		const unsigned cls = vxbRegularCellClass[cr.code];
		const int nverts = vxbRegularCellData[cls * 16] >> 4, ntri = vxbRegularCellData[cls * 16] & 0xF;
		const int local[3] = { lx, ly, lz };
		const int base[3] = { (bx * 16 + lx) * m, (by * 16 + ly) * m, (bz * 16 + lz) * m };
		uint32_t vid[12];
		cr.vbase = vcount;
		for (int k = 0; k < nverts; ++k)
		{
			const VxbVertexDesc d = vxb_regular_vertex_desc(vxbRegularVertexData[cr.code * 12 + k], cr.v);
			bool isNew = true, quirkV0 = false;
			if (!d.atC7 && (d.dir & mask) == d.dir) // dir == 8 never passes: mask < 8
			{
				const int oc = c - (d.dir & 1) - ((d.dir >> 1) & 1) * 16 - ((d.dir >> 2) & 1) * 256;
				const CellRec& o = cells[oc];
				const int ok = o.nontrivial ? ((o.slotK >> (4 * d.slot)) & 0xF) : VXB_NO_SLOT;
				if (ok != VXB_NO_SLOT)
				{
					if (o.matId == cr.matId)
					{
						isNew = false;
						vid[k] = o.vbase + __builtin_popcount(o.newMask & ((1u << ok) - 1));
					}
				}
				else if (d.endpoint) quirkV0 = true; // :1633-1640 creates the vertex at v0
				// interior + empty slot is unreachable (:1628-1631)
			}
			if (!isNew) continue;
			cr.newMask |= (uint16_t)(1u << k);
			vid[k] = vcount++;
			VxbRawVertex rv;
			if (d.endpoint) vxb_corner_vertex(g, level, base, local, quirkV0 ? d.v0 : ((d.t == 0) ? d.v1 : d.v0), cr.matId, cr.matBlend, rv);
			else vxb_edge_vertex(g, level, base, local, d, cr.matId, cr.matBlend, rv);
			vxb_regular_secondary(level, rv);
			raw.push_back(rv);
		}
		for (int tr = 0; tr < ntri * 3; ++tr) rawIdx.push_back(vid[vxbRegularCellData[cls * 16 + 1 + tr]]);
	}

	// degenerate filter + output conversion (PushBlocksToResult :1266-1369)
	if (!raw.empty())
	{
		ob.verts.resize(raw.size());
		for (size_t i = 0; i < raw.size(); ++i) { vxb_finish_vertex(raw[i], r.lut, ob.verts[i]); if (!r.lut.valid[raw[i].matId]) ob.verts[i].tex[0] = 0; } // unmapped material: textures stay zero (:1364-1368)
		for (size_t i = 0; i + 2 < rawIdx.size(); i += 3)
		{
			if (vxb_triangle_kept(raw[rawIdx[i]].p, raw[rawIdx[i + 1]].p, raw[rawIdx[i + 2]].p))
			{ ob.idx.push_back(rawIdx[i]); ob.idx.push_back(rawIdx[i + 1]); ob.idx.push_back(rawIdx[i + 2]); }
			else ++r.stats[3];
		}
	}

	if (!withTransitions) return;

	// ---- transition cells (:1754-2131) ----
	for (int face = 0; face < 6; ++face)
	{
		int axis, ua, va;
		vxb_face_axes(face, axis, ua, va);
		const int bc[3] = { bx, by, bz };
		if (face < 3 ? bc[axis] == 0 : bc[axis] == nb - 1) continue; // no neighbour block inside the grid (:1829-1835)

		struct TCell { uint8_t nontrivial, matId; uint16_t newMask; uint32_t vbase; uint8_t slotK[10]; };
		std::vector<TCell> tc(256);
		std::vector<VxbRawVertex> traw;
		uint32_t tcount = 0;
		for (int row = 0; row < 16; ++row)
		for (int col = 0; col < 16; ++col)
		{
			int local[3];
			local[axis] = (face >= 3) ? 15 : 0; local[ua] = col; local[va] = row;
			const int c = local[2] * 256 + local[1] * 16 + local[0];
			const int base[3] = { (bx * 16 + local[0]) * m, (by * 16 + local[1]) * m, (bz * 16 + local[2]) * m };
			TCell& t = tc[row * 16 + col];
			t.nontrivial = 0; t.newMask = 0; t.vbase = tcount; memset(t.slotK, VXB_NO_SLOT, sizeof(t.slotK));

			// material of the low-res cell; side effect on the cache for EVERY face cell (:1859)
			unsigned matId = VXB_EMPTY_MATERIAL, matBlend = 0;
			if (voteCell(r, level, base, matId, matBlend)) cacheStore(r, level, bid, c, matId, matBlend);
			else { matId = VXB_EMPTY_MATERIAL; matBlend = 0; }
			t.matId = (uint8_t)matId;

			signed char v[13];
			for (int i = 0; i < 13; ++i)
			{
				int p[3];
				vxb_transition_sample_pos(face, level, base, i, p);
				v[i] = (signed char)vxb_dist(g, p[0], p[1], p[2]);
			}
			const unsigned code = vxb_transition_case_code(v);
			if (code == 0 || code == 511) continue;
			t.nontrivial = 1;
			const unsigned cls = vxbTransitionCellClass[code];
			const unsigned char* cd = &vxbTransitionCellData[(cls & 0x7F) * 40];
			const int nverts = cd[0] >> 4, ntri = cd[0] & 0xF;

			int mask = (row > 0) ? 2 : 0;
			for (int x = 0; x < col; ++x) if (tc[row * 16 + x].nontrivial) mask |= 1;

			uint32_t vid[12];
			for (int k = 0; k < nverts; ++k)
			{
				const VxbTransVertexDesc d = vxb_transition_vertex_desc(vxbTransitionVertexData[code * 12 + k], v, vxbTransitionCornerData);
				bool isNew = true;
				if ((d.dir & mask) == d.dir)
				{
					const TCell& o = tc[(row - ((d.dir >> 1) & 1)) * 16 + (col - (d.dir & 1))];
					const int ok = o.nontrivial ? o.slotK[d.slot] : VXB_NO_SLOT;
					if (ok != VXB_NO_SLOT && o.matId == t.matId)
					{
						isNew = false;
						vid[k] = o.vbase + __builtin_popcount(o.newMask & ((1u << ok) - 1));
					}
				}
				else if (d.dir == 8) t.slotK[d.slot] = (uint8_t)k; // stored only when no reuse was attempted (:2097)
				if (!isNew) continue;
				t.newMask |= (uint16_t)(1u << k);
				vid[k] = tcount++;
				VxbRawVertex rv;
				vxb_transition_vertex(g, face, level, base, local, d, matId, matBlend, rv);
				traw.push_back(rv);
			}
			const bool flip = ((cls >> 7) & 1) ^ (face & 1);
			for (int tr = 0; tr < ntri; ++tr)
			{
				const uint32_t a = vid[cd[1 + tr * 3]], b = vid[cd[2 + tr * 3]], cc = vid[cd[3 + tr * 3]];
				ob.tidx[face].push_back(a);
				ob.tidx[face].push_back(flip ? cc : b);
				ob.tidx[face].push_back(flip ? b : cc);
			}
		}
		ob.tverts[face].resize(traw.size());
		for (size_t i = 0; i < traw.size(); ++i) { vxb_finish_vertex(traw[i], r.lut, ob.tverts[face][i]); if (!r.lut.valid[traw[i].matId]) ob.tverts[face][i].tex[0] = 0; }
	}
}

bool blockSkipped(const Run& r, int bx, int by, int bz)
{
	const int nb = r.n / 16;
	for (int z = -1; z < 2; ++z) for (int y = -1; y < 2; ++y) for (int x = -1; x < 2; ++x)
	{
		const int cx = vxb_clampi(bx + x, 0, nb - 1), cy = vxb_clampi(by + y, 0, nb - 1), cz = vxb_clampi(bz + z, 0, nb - 1);
		if (!r.emptyFlags[coordId(nb, cx, cy, cz)]) return false;
	}
	return true;
}

} // namespace

extern "C"
{

// materialTable: 256 x 6 bytes or null (identity); validMask: 256 bytes or null
void* vxr_run(unsigned n, const signed char* dist, const unsigned char* mat, const unsigned char* blend,
	const unsigned char* materialTable, const unsigned char* validMask, int maxLevels)
{
	Run* r = new Run;
	r->grid.dist = dist; r->grid.mat = mat; r->grid.blend = blend; r->grid.n = (int)n;
	r->n = (int)n;
	for (unsigned i = 0; i < 256; ++i)
	{
		unsigned d0[3], d1[3];
		for (int k = 0; k < 3; ++k) { d0[k] = materialTable ? materialTable[i * 6 + k] : i; d1[k] = materialTable ? materialTable[i * 6 + 3 + k] : i; }
		r->lut.tex0[i] = (d1[1] << 16) | (d0[1] << 24);
		r->lut.tex1[i] = d1[2] | (d1[0] << 8) | (d0[2] << 16) | (d0[0] << 24);
		r->lut.valid[i] = validMask ? (validMask[i] != 0) : 1;
	}
	int levels = 1;
	for (unsigned v = n >> 4; v >>= 1;) ++levels;
	r->levels = levels;
	memset(r->stats, 0, sizeof(r->stats));
	r->nextId = 0;
	r->cache.resize(levels);
	r->out.resize(levels);

	const int nb0 = (int)n / 16;
	r->emptyFlags.resize((size_t)nb0 * nb0 * nb0);
	for (int bz = 0; bz < nb0; ++bz) for (int by = 0; by < nb0; ++by) for (int bx = 0; bx < nb0; ++bx)
		r->emptyFlags[coordId(nb0, bx, by, bz)] = blockIsEmpty(r->grid, bx, by, bz);

	for (int level = 0; level < levels; ++level)
	{
		const int m = 1 << level, nb = (int)n / 16 / m;
		const bool compute = (maxLevels <= 0 || level < maxLevels);
		for (int bz = 0; bz < nb; ++bz) for (int by = 0; by < nb; ++by) for (int bx = 0; bx < nb; ++bx)
		{
			const uint32_t id = r->nextId++;
			if (!compute) continue;
			++r->stats[0];
			if (level == 0 && blockSkipped(*r, bx, by, bz)) continue;
			OutBlock ob;
			ob.id = id; ob.coordId = coordId(nb, bx, by, bz);
			processBlock(*r, level, bx, by, bz, level > 0 && level != levels - 1, ob);
			if (ob.verts.empty()) continue;
			const float e = (float)(16 * m);
			ob.mn[0] = bx * e; ob.mn[1] = bz * e; ob.mn[2] = by * e; // y/z swapped on output (:1289-1291)
			ob.mx[0] = ob.mn[0] + e; ob.mx[1] = ob.mn[1] + e; ob.mx[2] = ob.mn[2] + e;
			r->out[level].push_back(std::move(ob));
		}
	}
	return r;
}

void vxr_destroy(void* h) { delete static_cast<Run*>(h); }
unsigned vxr_levels(void* h) { return (unsigned)static_cast<Run*>(h)->levels; }
unsigned vxr_blocks(void* h, unsigned level) { return (unsigned)static_cast<Run*>(h)->out[level].size(); }
void vxr_stats(void* h, unsigned* out20) { memcpy(out20, static_cast<Run*>(h)->stats, sizeof(unsigned) * 20); }
void vxr_empty_flags(void* h, unsigned char* out) { Run* r = static_cast<Run*>(h); memcpy(out, r->emptyFlags.data(), r->emptyFlags.size()); }

void vxr_level_totals(void* h, unsigned level, uint64_t* t4)
{
	t4[0] = t4[1] = t4[2] = t4[3] = 0;
	for (const OutBlock& b : static_cast<Run*>(h)->out[level])
	{
		t4[0] += b.verts.size(); t4[1] += b.idx.size();
		for (int f = 0; f < 6; ++f) { t4[2] += b.tverts[f].size(); t4[3] += b.tidx[f].size(); }
	}
}

// same flattened layout as tests/harness/vxh_capi.cpp: vxh_surface_level_dump
void vxr_level_dump(void* h, unsigned level, void* rowsOut, void* vertices, unsigned* indices, void* transVertices, unsigned* transIndices)
{
	struct Row { uint32_t id; float mn[3], mx[3]; uint32_t nv, ni, tnv[6], tni[6]; };
	Row* rows = static_cast<Row*>(rowsOut);
	VxbVertex* v = static_cast<VxbVertex*>(vertices);
	VxbVertex* tv = static_cast<VxbVertex*>(transVertices);
	size_t i = 0;
	for (const OutBlock& b : static_cast<Run*>(h)->out[level])
	{
		Row& row = rows[i++];
		row.id = b.id;
		memcpy(row.mn, b.mn, 12); memcpy(row.mx, b.mx, 12);
		row.nv = (uint32_t)b.verts.size(); row.ni = (uint32_t)b.idx.size();
		if (row.nv) { memcpy(v, b.verts.data(), sizeof(VxbVertex) * row.nv); v += row.nv; }
		if (row.ni) { memcpy(indices, b.idx.data(), 4 * (size_t)row.ni); indices += row.ni; }
		for (int f = 0; f < 6; ++f)
		{
			row.tnv[f] = (uint32_t)b.tverts[f].size(); row.tni[f] = (uint32_t)b.tidx[f].size();
			if (row.tnv[f]) { memcpy(tv, b.tverts[f].data(), sizeof(VxbVertex) * row.tnv[f]); tv += row.tnv[f]; }
			if (row.tni[f]) { memcpy(transIndices, b.tidx[f].data(), 4 * (size_t)row.tni[f]); transIndices += row.tni[f]; }
		}
	}
}

} // extern "C"
