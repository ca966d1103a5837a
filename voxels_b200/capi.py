"""ctypes binding of include/vxb200.h (libvxb200.so).  Mirrors the C ABI one to one."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
FLAG_NO_TRANSITIONS = 1
FLAG_KERNEL_TIMES = 2

VERTEX_DTYPE = np.dtype([("pos", "<f4", 3), ("sec", "<f4", 4), ("nrm", "<f4", 3), ("tex", "u1", 8)])
RECORD_DTYPE = np.dtype([("level", "<u4"), ("coord_id", "<u4"), ("id", "<u4"), ("vertex_count", "<u4"),
                         ("index_count", "<u4"), ("vertex_offset", "<u4"), ("index_offset", "<u4"),
                         ("trans_vertex_count", "<u4", 6), ("trans_index_count", "<u4", 6),
                         ("trans_vertex_offset", "<u4", 6), ("trans_index_offset", "<u4", 6), ("reserved", "<u4")])
assert VERTEX_DTYPE.itemsize == 48 and RECORD_DTYPE.itemsize == 128

# what tests/harness.py calls a block row (the reference-facing view of a block)
ROW_DTYPE = np.dtype([("id", "<u4"), ("min", "<f4", 3), ("max", "<f4", 3), ("nv", "<u4"), ("ni", "<u4"),
                      ("tnv", "<u4", 6), ("tni", "<u4", 6)])


class ResultInfo(C.Structure):
    _fields_ = [("levels_total", C.c_uint32), ("levels_computed", C.c_uint32), ("block_count", C.c_uint32),
                ("pad", C.c_uint32), ("vertex_span", C.c_uint64), ("index_span", C.c_uint64),
                ("trans_vertex_span", C.c_uint64), ("trans_index_span", C.c_uint64), ("vertex_total", C.c_uint64),
                ("index_total", C.c_uint64), ("trans_vertex_total", C.c_uint64), ("trans_index_total", C.c_uint64),
                ("stats", C.c_uint32 * 20), ("used_materials", C.c_uint32 * 8), ("device_ms", C.c_float),
                ("kernel_launches", C.c_uint32)]


class RegionInfo(C.Structure):
    _fields_ = [("levels", C.c_uint32), ("pad", C.c_uint32), ("min_dirty", (C.c_float * 3) * 12), ("max_dirty", (C.c_float * 3) * 12),
                ("id_start", C.c_uint32 * 12), ("block_count", C.c_uint32 * 12)]


class ShardBuffers(C.Structure):
    _fields_ = [("block_info", C.c_void_p), ("block_info_bytes", C.c_uint64), ("chunk_bytes", C.c_uint64), ("pages", C.c_void_p),
                ("pages_bytes", C.c_uint64), ("valid", C.c_void_p), ("valid_bytes", C.c_uint64), ("super_level", C.c_uint32),
                ("pad", C.c_uint32)]


class Surface(C.Structure):
    """vxb_surface (include/vxb200.h): a built-in procedural surface."""
    _fields_ = [("kind", C.c_uint32), ("material", C.c_uint32), ("blend", C.c_uint32), ("seed", C.c_uint32), ("p", C.c_float * 8)]

    @staticmethod
    def sphere(center, radius, material=0, blend=0):
        s = Surface(); s.kind = 0; s.material = material; s.blend = blend
        s.p[0], s.p[1], s.p[2], s.p[3] = center[0], center[1], center[2], radius
        return s

    @staticmethod
    def plane(normal, d0, material=0, blend=0):
        s = Surface(); s.kind = 1; s.material = material; s.blend = blend
        s.p[0], s.p[1], s.p[2], s.p[3] = normal[0], normal[1], normal[2], d0
        return s

    @staticmethod
    def terrain(period, origin=(0.0, 0.0), seed=1234):
        s = Surface(); s.kind = 2; s.seed = seed
        s.p[0], s.p[1], s.p[2] = float(period), float(origin[0]), float(origin[1])
        return s


class DrawLists(C.Structure):
    _fields_ = [("regular_count", C.c_uint32), ("transition_count", C.c_uint32), ("counts", C.c_void_p), ("regular", C.c_void_p),
                ("regular_info", C.c_void_p), ("transition", C.c_void_p), ("transition_info", C.c_void_p)]


DRAW_COMMAND_DTYPE = np.dtype([("index_count", "<u4"), ("instance_count", "<u4"), ("first_index", "<u4"), ("base_vertex", "<i4"), ("first_instance", "<u4")])
DRAW_INFO_DTYPE = np.dtype([("block_id", "<u4"), ("level", "<u4"), ("block_adj", "<u4"), ("face", "<u4")])


class NcclId(C.Structure):
    _fields_ = [("internal", C.c_char * 128)]


ERR_CAPACITY = -4

# vxb_pack_producer (include/vxb200.h): void (*)(void* user, uint32_t layer0, uint32_t layer1)
PackProducer = C.CFUNCTYPE(None, C.c_void_p, C.c_uint32, C.c_uint32)


class DevicePointer:
    """A raw device range as a CUDA-array-interface object: torch.as_tensor(DevicePointer(p, nbytes), device=...) is a
    zero-copy uint8 view (how dist.py hands the exchange buffers of a sharded run to torch.distributed)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 3,
                                         "strides": None}


class VxbError(RuntimeError):
    pass


def library_path():
    return os.path.join(_HERE, "lib", "libvxb200.so")


_lib = None


def load_library():
    """Loads libvxb200.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise VxbError("%s is missing - run `python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc); "
                       "there is no CPU fallback" % path)
    L = C.CDLL(path)
    vp, u32, u64 = C.c_void_p, C.c_uint32, C.c_uint64
    sig = {
        "vxb_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
        "vxb_destroy": (None, [vp]),
        "vxb_last_error": (C.c_char_p, [vp]),
        "vxb_stream": (vp, [vp]),
        "vxb_exchange_stream": (vp, [vp]),
        "vxb_grid_upload_dense": (C.c_int, [vp, u32, vp, vp, vp]),
        "vxb_grid_upload_blocks": (C.c_int, [vp, u32, vp, vp, vp]),
        "vxb_grid_set_device": (C.c_int, [vp, u32, vp, vp, vp]),
        "vxb_grid_upload_packed": (C.c_int, [vp, vp, C.c_size_t]),
        "vxb_grid_upload_packed_streamed": (C.c_int, [vp, vp, C.c_size_t, PackProducer, vp]),
        "vxb_result_download_begin": (C.c_int, [vp, vp, vp, vp, vp, vp]),
        "vxb_result_download_end": (C.c_int, [vp]),
        "vxb_pack_dense_bound": (C.c_size_t, [u32]),
        "vxb_pack_dense": (C.c_int, [u32, vp, vp, vp, vp, C.c_size_t, C.POINTER(C.c_size_t)]),
        "vxb_grid_device_pointers": (C.c_int, [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]),
        "vxb_set_materials": (C.c_int, [vp, vp, vp]),
        "vxb_grid_fill": (C.c_int, [vp, u32, C.POINTER(Surface), vp, C.c_float]),
        "vxb_grid_inject_surface": (C.c_int, [vp, vp, vp, C.POINTER(Surface), C.c_int, vp, vp]),
        "vxb_grid_inject_material": (C.c_int, [vp, vp, vp, u32, C.c_int, vp, vp]),
        "vxb_grid_pack": (C.c_int, [vp, vp, C.c_size_t, C.POINTER(C.c_size_t)]),
        "vxb_grid_download_dense": (C.c_int, [vp, vp, vp, vp]),
        "vxb_result_device_arenas": (C.c_int, [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]),
        "vxb_result_select_lod": (C.c_int, [vp, vp, C.c_float, C.POINTER(DrawLists)]),
        "vxb_result_download_draws": (C.c_int, [vp, vp, vp, vp, vp]),
        "vxb_polygonize": (C.c_int, [vp, u32, u32]),
        "vxb_result_info_get": (C.c_int, [vp, C.POINTER(ResultInfo)]),
        "vxb_grid_update_blocks": (C.c_int, [vp, u32, vp, vp, vp, vp]),
        "vxb_polygonize_region": (C.c_int, [vp, vp, vp, u32]),
        "vxb_region_info_get": (C.c_int, [vp, C.POINTER(RegionInfo)]),
        "vxb_polygonize_sharded": (C.c_int, [vp, u32, u32]),
        "vxb_shard_configure": (C.c_int, [vp, u32, u32, u32]),
        "vxb_shard_buffers_get": (C.c_int, [vp, C.POINTER(ShardBuffers)]),
        "vxb_shard_set_peer": (C.c_int, [vp, u32, vp, vp]),
        "vxb_shard_export": (C.c_int, [vp, C.POINTER(C.c_int)]),
        "vxb_shard_import": (C.c_int, [vp, u32, C.c_int]),
        "vxb_shard_nccl_unique_id": (C.c_int, [C.POINTER(NcclId)]),
        "vxb_shard_nccl_init": (C.c_int, [vp, C.POINTER(NcclId), u32, u32]),
        "vxb_cube_create": (C.c_int, [vp, u32, u32, u32, u32]),
        "vxb_cube_info": (C.c_int, [vp, C.POINTER(u32), C.POINTER(u32), C.POINTER(u64 * 4)]),
        "vxb_cube_export": (C.c_int, [vp, u32, u32, C.POINTER(C.c_int)]),
        "vxb_cube_import": (C.c_int, [vp, u32, u32, C.c_int]),
        "vxb_cube_piece": (C.c_int, [vp, u32, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(u64)]),
        "vxb_result_download": (C.c_int, [vp, vp, vp, vp, vp, vp]),
        "vxb_set_capacity": (C.c_int, [vp, u64, u64, u64, u64]),
        "vxb_kernel_ms": (C.c_int, [vp, C.c_int, C.POINTER(C.c_float), C.POINTER(u32)]),
        "vxb_result_unmapped_materials": (u64, [vp, vp, u64]),
        "vxb_host_alloc": (vp, [C.c_size_t]),
        "vxb_host_free": (None, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)  # AttributeError = a symbol include/vxb200.h declares is not exported
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


EXPORTED_SYMBOLS = ["vxb_create", "vxb_destroy", "vxb_last_error", "vxb_stream", "vxb_exchange_stream", "vxb_grid_upload_dense",
                    "vxb_grid_upload_blocks", "vxb_grid_upload_packed", "vxb_pack_dense_bound", "vxb_pack_dense", "vxb_grid_set_device", "vxb_grid_device_pointers", "vxb_set_materials",
                    "vxb_polygonize", "vxb_polygonize_region", "vxb_region_info_get", "vxb_grid_update_blocks", "vxb_result_info_get", "vxb_result_download", "vxb_set_capacity", "vxb_kernel_ms",
                    "vxb_result_unmapped_materials", "vxb_host_alloc", "vxb_host_free", "vxb_polygonize_sharded", "vxb_shard_configure",
                    "vxb_shard_buffers_get", "vxb_shard_set_peer", "vxb_shard_export", "vxb_shard_import", "vxb_shard_nccl_unique_id",
                    "vxb_shard_nccl_init", "vxb_grid_fill", "vxb_grid_inject_surface", "vxb_grid_inject_material", "vxb_grid_pack",
                    "vxb_grid_download_dense", "vxb_result_device_arenas", "vxb_result_select_lod", "vxb_result_download_draws", "vxb_cube_create", "vxb_cube_info", "vxb_cube_export", "vxb_cube_import", "vxb_cube_piece",
                    "vxb_grid_upload_packed_streamed", "vxb_result_download_begin", "vxb_result_download_end"]


def nccl_unique_id():
    """128 bytes from ncclGetUniqueId (through libvxb200.so's dlopen of libnccl.so.2)."""
    L = load_library()
    nid = NcclId()
    rc = L.vxb_shard_nccl_unique_id(C.byref(nid))
    if rc != 0:
        raise VxbError("vxb_shard_nccl_unique_id failed (%d): %s" % (rc, L.vxb_last_error(None).decode()))
    return bytes(C.string_at(C.byref(nid), 128))


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return a.ctypes.data_as(C.c_void_p)


def pack_dense(dist, mat, blend, out=None):
    """Host helper: the reference's PackForSave bytes for dense [z,y,x] volumes (numpy uint8 array).  No GPU needed."""
    L = load_library()
    n = dist.shape[0]
    bound = L.vxb_pack_dense_bound(n)
    buf = np.empty(bound, np.uint8) if out is None else out
    written = C.c_size_t(0)
    rc = L.vxb_pack_dense(n, _ptr(dist), _ptr(mat), _ptr(blend), _ptr(buf), buf.size, C.byref(written))
    if rc != 0:
        raise VxbError("vxb_pack_dense failed (%d)" % rc)
    return buf[:written.value]


class LevelView:
    """One LOD level in the layout tests/harness.LevelDump uses (rows + concatenated arrays, block order)."""

    def __init__(self, rows, verts, idx, tverts, tidx):
        self.rows, self.verts, self.idx, self.tverts, self.tidx = rows, verts, idx, tverts, tidx


class Result:
    def __init__(self, n, info, records, verts, idx, tverts, tidx):
        self.n, self.info, self.records = n, info, records
        self.verts, self.idx, self.tverts, self.tidx = verts, idx, tverts, tidx
        self.stats = np.array(list(info.stats), np.uint32)

    def level(self, level):
        recs = self.records[self.records["level"] == level]
        m = 16 << level
        nb = self.n // m
        rows = np.zeros(len(recs), ROW_DTYPE)
        cid = recs["coord_id"].astype(np.int64)
        bx, by, bz = cid % nb, (cid // nb) % nb, cid // (nb * nb)
        rows["id"] = recs["id"]
        rows["min"] = np.stack([bx, bz, by], axis=1).astype(np.float32) * m  # y/z swapped on output
        rows["max"] = rows["min"] + np.float32(m)
        rows["nv"], rows["ni"] = recs["vertex_count"], recs["index_count"]
        rows["tnv"], rows["tni"] = recs["trans_vertex_count"], recs["trans_index_count"]

        def gather(arr, offs, counts):
            parts = [arr[o:o + c] for o, c in zip(offs, counts) if c]
            return np.concatenate(parts) if parts else arr[:0]

        verts = gather(self.verts, recs["vertex_offset"], recs["vertex_count"])
        idx = gather(self.idx, recs["index_offset"], recs["index_count"])
        tverts = gather(self.tverts, recs["trans_vertex_offset"].ravel(), recs["trans_vertex_count"].ravel())
        tidx = gather(self.tidx, recs["trans_index_offset"].ravel(), recs["trans_index_count"].ravel())
        return LevelView(rows, verts, idx, tverts, tidx)


def merge_results(results):
    """The per-rank results of a sharded run as ONE result: arenas concatenated in rank order, directory offsets
    rebased and sorted by (level, coord_id) - the reference's block order - statistics summed.  Block ids need no
    fixing: they are a function of (level, coordinate) in a full run (TransVoxelImpl.cpp:395-401)."""
    recs, bases = [], [0, 0, 0, 0]
    for r in results:
        rr = r.records.copy()
        rr["vertex_offset"] += np.uint32(bases[0]); rr["index_offset"] += np.uint32(bases[1])
        rr["trans_vertex_offset"] += np.uint32(bases[2]); rr["trans_index_offset"] += np.uint32(bases[3])
        recs.append(rr)
        bases = [bases[0] + len(r.verts), bases[1] + len(r.idx), bases[2] + len(r.tverts), bases[3] + len(r.tidx)]
    records = np.concatenate(recs)
    records = records[np.lexsort((records["coord_id"], records["level"]))]
    merged = Result.__new__(Result)
    merged.n, merged.info, merged.records = results[0].n, None, records
    merged.verts = np.concatenate([r.verts for r in results]); merged.idx = np.concatenate([r.idx for r in results])
    merged.tverts = np.concatenate([r.tverts for r in results]); merged.tidx = np.concatenate([r.tidx for r in results])
    merged.stats = np.sum([r.stats.astype(np.uint64) for r in results], axis=0).astype(np.uint32)
    return merged


class Context:
    """vxb_context wrapper.  Raises VxbError on any failure (no CUDA device included)."""

    def __init__(self, device=0):
        self.L = load_library()
        h = C.c_void_p()
        rc = self.L.vxb_create(device, C.byref(h))
        if rc != 0:
            raise VxbError("vxb_create failed (%d): %s" % (rc, self.L.vxb_last_error(None).decode()))
        self.h = h
        self.n = 0
        self._keep = None

    def close(self):
        if getattr(self, "h", None):
            self.L.vxb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise VxbError("%s failed (%d): %s" % (what, rc, self.L.vxb_last_error(self.h).decode()))

    def upload_dense(self, dist, mat=None, blend=None):
        """dist/mat/blend: host arrays (numpy or anything with .ctypes / an int address) indexed [z, y, x]."""
        n = dist.shape[0]
        self._check(self.L.vxb_grid_upload_dense(self.h, n, _ptr(dist), _ptr(mat), _ptr(blend)), "vxb_grid_upload_dense")
        self.n = n

    def upload_dense_ptr(self, n, dist_ptr, mat_ptr, blend_ptr):
        self._check(self.L.vxb_grid_upload_dense(self.h, n, C.c_void_p(dist_ptr), C.c_void_p(mat_ptr), C.c_void_p(blend_ptr)),
                    "vxb_grid_upload_dense")
        self.n = n

    def upload_blocks(self, n, dist_blocks, mat_blocks=None, blend_blocks=None):
        self._check(self.L.vxb_grid_upload_blocks(self.h, n, _ptr(dist_blocks), _ptr(mat_blocks), _ptr(blend_blocks)),
                    "vxb_grid_upload_blocks")
        self.n = n

    def upload_packed(self, blob, size=None):
        """blob: the bytes of Grid::PackForSave (numpy uint8 array, bytes, or an int address with `size`)."""
        if isinstance(blob, int):
            ptr, nbytes = C.c_void_p(blob), size
        elif isinstance(blob, (bytes, bytearray)):
            arr = np.frombuffer(blob, np.uint8); ptr, nbytes = _ptr(arr), arr.size
            self._blob_keep = arr
        else:
            ptr, nbytes = _ptr(blob), blob.size
        self._check(self.L.vxb_grid_upload_packed(self.h, ptr, nbytes), "vxb_grid_upload_packed")
        self.n = int(np.frombuffer(C.string_at(ptr, 8), np.uint32)[1])

    def upload_packed_streamed(self, blob, produce):
        """vxb_grid_upload_packed_streamed: `blob` (numpy uint8) holds the header and the complete size table; produce(layer0,
        layer1) is called on this thread before the data of those block layers is copied and must have written it."""
        cb = PackProducer(lambda user, l0, l1: produce(int(l0), int(l1)))
        self._check(self.L.vxb_grid_upload_packed_streamed(self.h, _ptr(blob), blob.size, cb, None), "vxb_grid_upload_packed_streamed")
        self.n = int(blob[:8].view(np.uint32)[1])

    # ---- device-resident grid store (fill / edit / pack) ----
    def fill(self, n, surface, start=(0.0, 0.0, 0.0), step=1.0):
        """Grid::Create(n, n, n, start, step, &surface) on the device (a capi.Surface); with a cube: this rank's pieces."""
        st = np.ascontiguousarray(start, np.float32)
        self._check(self.L.vxb_grid_fill(self.h, n, C.byref(surface), _ptr(st), step), "vxb_grid_fill")
        self.n = n

    def inject_surface(self, position, extents, surface, inject_type):
        """VoxelGrid::InjectSurface on the device grid; returns the modified box (6 floats, Y-up) the reference returns."""
        p = np.ascontiguousarray(position, np.float32); e = np.ascontiguousarray(extents, np.float32)
        lo = np.zeros(3, np.float32); hi = np.zeros(3, np.float32)
        self._check(self.L.vxb_grid_inject_surface(self.h, _ptr(p), _ptr(e), C.byref(surface), inject_type, _ptr(lo), _ptr(hi)), "vxb_grid_inject_surface")
        return np.concatenate([lo, hi])

    def inject_material(self, position, extents, material, add_subtract_blend):
        p = np.ascontiguousarray(position, np.float32); e = np.ascontiguousarray(extents, np.float32)
        lo = np.zeros(3, np.float32); hi = np.zeros(3, np.float32)
        self._check(self.L.vxb_grid_inject_material(self.h, _ptr(p), _ptr(e), material, 1 if add_subtract_blend else 0, _ptr(lo), _ptr(hi)), "vxb_grid_inject_material")
        return np.concatenate([lo, hi])

    def pack(self):
        """Grid::PackForSave of the device grid (GPU run-length coding): numpy uint8 blob."""
        buf = np.empty(self.L.vxb_pack_dense_bound(self.n), np.uint8)
        written = C.c_size_t(0)
        self._check(self.L.vxb_grid_pack(self.h, _ptr(buf), buf.size, C.byref(written)), "vxb_grid_pack")
        return buf[:written.value].copy()

    def download_dense(self):
        n = self.n
        dist = np.empty((n, n, n), np.int8); mat = np.empty((n, n, n), np.uint8); blend = np.empty((n, n, n), np.uint8)
        self._check(self.L.vxb_grid_download_dense(self.h, _ptr(dist), _ptr(mat), _ptr(blend)), "vxb_grid_download_dense")
        return dist, mat, blend

    # ---- consumer side: LOD cut + indirect draw lists on the device ----
    def select_lod(self, camera, base_distance):
        """(regular commands, regular infos, transition commands, transition infos) as numpy record arrays (copied back for
        inspection; a renderer consumes the device-resident lists of vxb_draw_lists directly)."""
        cam = np.ascontiguousarray(camera, np.float32)
        lists = DrawLists()
        self._check(self.L.vxb_result_select_lod(self.h, _ptr(cam), base_distance, C.byref(lists)), "vxb_result_select_lod")
        rc = np.zeros(lists.regular_count, DRAW_COMMAND_DTYPE); ri = np.zeros(lists.regular_count, DRAW_INFO_DTYPE)
        tc = np.zeros(lists.transition_count, DRAW_COMMAND_DTYPE); ti = np.zeros(lists.transition_count, DRAW_INFO_DTYPE)
        self._check(self.L.vxb_result_download_draws(self.h, _ptr(rc), _ptr(ri), _ptr(tc), _ptr(ti)), "vxb_result_download_draws")
        return rc, ri, tc, ti

    def device_arenas(self):
        v, i, tv, ti = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._check(self.L.vxb_result_device_arenas(self.h, C.byref(v), C.byref(i), C.byref(tv), C.byref(ti)), "vxb_result_device_arenas")
        return v.value, i.value, tv.value, ti.value

    def set_device_grid(self, n, d_dist, d_mat, d_blend, keep=None):
        """Device pointers (ints) of dense volumes that stay resident (e.g. torch tensors' data_ptr())."""
        self._check(self.L.vxb_grid_set_device(self.h, n, C.c_void_p(d_dist), C.c_void_p(d_mat), C.c_void_p(d_blend)),
                    "vxb_grid_set_device")
        self.n = n
        self._keep = keep

    def set_materials(self, table=None, valid=None):
        self._check(self.L.vxb_set_materials(self.h, _ptr(table), _ptr(valid)), "vxb_set_materials")

    def set_capacity(self, vertices=0, indices=0, trans_vertices=0, trans_indices=0):
        self._check(self.L.vxb_set_capacity(self.h, vertices, indices, trans_vertices, trans_indices), "vxb_set_capacity")

    def polygonize(self, max_levels=0, flags=0):
        self._check(self.L.vxb_polygonize(self.h, max_levels, flags), "vxb_polygonize")
        return self.info()

    def update_blocks(self, coords, dist_blocks, mat_blocks=None, blend_blocks=None):
        """coords: uint32 [count, 3] block coordinates (x, y, z); *_blocks: uint8/int8 [count, 4096]."""
        coords = np.ascontiguousarray(coords, np.uint32)
        self._check(self.L.vxb_grid_update_blocks(self.h, len(coords), _ptr(coords), _ptr(dist_blocks), _ptr(mat_blocks), _ptr(blend_blocks)),
                    "vxb_grid_update_blocks")

    def polygonize_region(self, min_corner, max_corner, flags=0):
        lo = np.ascontiguousarray(min_corner, np.float32); hi = np.ascontiguousarray(max_corner, np.float32)
        self._check(self.L.vxb_polygonize_region(self.h, _ptr(lo), _ptr(hi), flags), "vxb_polygonize_region")
        return self.info()

    # ---- sharded runs (include/vxb200.h: vxb_polygonize_sharded, vxb_shard_*, vxb_cube_*) ----
    def shard_configure(self, rank, world, group_planes=0):
        self._check(self.L.vxb_shard_configure(self.h, rank, world, group_planes), "vxb_shard_configure")

    def shard_buffers(self):
        x = ShardBuffers()
        self._check(self.L.vxb_shard_buffers_get(self.h, C.byref(x)), "vxb_shard_buffers_get")
        return x

    def shard_set_peer(self, peer, pages, valid):
        self._check(self.L.vxb_shard_set_peer(self.h, peer, C.c_void_p(pages), C.c_void_p(valid)), "vxb_shard_set_peer")

    def shard_export(self):
        fd = C.c_int(-1)
        self._check(self.L.vxb_shard_export(self.h, C.byref(fd)), "vxb_shard_export")
        return fd.value

    def shard_import(self, peer, fd):
        self._check(self.L.vxb_shard_import(self.h, peer, fd), "vxb_shard_import")

    def shard_nccl_init(self, id_bytes, rank, world):
        nid = NcclId()
        C.memmove(C.byref(nid), bytes(id_bytes), 128)
        self._check(self.L.vxb_shard_nccl_init(self.h, C.byref(nid), rank, world), "vxb_shard_nccl_init")

    def polygonize_sharded(self, phase, flags=0):
        """phase 0 / 1 / 2: the pieces around the two exchanges; 3: the whole step (NCCL inside).  Returns 0, or
        ERR_CAPACITY from phase 2 / 3 (arenas grown: every rank repeats the run)."""
        rc = self.L.vxb_polygonize_sharded(self.h, phase, flags)
        if rc == ERR_CAPACITY and phase >= 2:
            return rc
        self._check(rc, "vxb_polygonize_sharded")
        return 0

    def cube_create(self, n, rank, world, group_planes=0):
        self._check(self.L.vxb_cube_create(self.h, n, rank, world, group_planes), "vxb_cube_create")
        self.n = n

    def cube_info(self):
        """(pieces, channels, bytes per piece of each channel)."""
        pieces, channels, sizes = C.c_uint32(0), C.c_uint32(0), (C.c_uint64 * 4)()
        self._check(self.L.vxb_cube_info(self.h, C.byref(pieces), C.byref(channels), C.byref(sizes)), "vxb_cube_info")
        return pieces.value, channels.value, list(sizes)

    def cube_export(self, channel, piece):
        fd = C.c_int(-1)
        self._check(self.L.vxb_cube_export(self.h, channel, piece, C.byref(fd)), "vxb_cube_export")
        return fd.value

    def cube_import(self, channel, piece, fd):
        self._check(self.L.vxb_cube_import(self.h, channel, piece, fd), "vxb_cube_import")

    def cube_piece(self, piece):
        """(dist, mat, blend) device addresses of a piece and its size in bytes per channel."""
        d, m, b, size = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_uint64(0)
        self._check(self.L.vxb_cube_piece(self.h, piece, C.byref(d), C.byref(m), C.byref(b), C.byref(size)), "vxb_cube_piece")
        return d.value, m.value, b.value, size.value

    def device_pointers(self):
        d, m, b = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._check(self.L.vxb_grid_device_pointers(self.h, C.byref(d), C.byref(m), C.byref(b)), "vxb_grid_device_pointers")
        return d.value, m.value, b.value

    def stream(self):
        return self.L.vxb_stream(self.h)

    def exchange_stream(self):
        return self.L.vxb_exchange_stream(self.h)

    def region_info(self):
        ri = RegionInfo()
        self._check(self.L.vxb_region_info_get(self.h, C.byref(ri)), "vxb_region_info_get")
        return ri

    def info(self):
        info = ResultInfo()
        self._check(self.L.vxb_result_info_get(self.h, C.byref(info)), "vxb_result_info_get")
        return info

    def kernel_ms(self, which):
        ms, launches = C.c_float(0), C.c_uint32(0)
        self._check(self.L.vxb_kernel_ms(self.h, which, C.byref(ms), C.byref(launches)), "vxb_kernel_ms")
        return ms.value, launches.value

    def unmapped_materials(self):
        n = self.L.vxb_result_unmapped_materials(self.h, None, 0)
        out = np.zeros(n, np.uint8)
        if n:
            self.L.vxb_result_unmapped_materials(self.h, _ptr(out), n)
        return out

    def download_begin(self):
        """vxb_result_download_begin into fresh numpy arrays: returns a Result whose records are valid at once and whose
        arenas are valid after download_end()."""
        info = self.info()
        records = np.zeros(info.block_count, RECORD_DTYPE)
        verts = np.zeros(info.vertex_span, VERTEX_DTYPE); idx = np.zeros(info.index_span, np.uint32)
        tverts = np.zeros(info.trans_vertex_span, VERTEX_DTYPE); tidx = np.zeros(info.trans_index_span, np.uint32)
        self._check(self.L.vxb_result_download_begin(self.h, _ptr(records), _ptr(verts), _ptr(idx), _ptr(tverts), _ptr(tidx)),
                    "vxb_result_download_begin")
        return Result(self.n, info, records, verts, idx, tverts, tidx)

    def download_end(self):
        self._check(self.L.vxb_result_download_end(self.h), "vxb_result_download_end")

    def download(self, into=None):
        """Device -> host copy of the directory and the arenas.  `into` may supply preallocated (pinned) buffers:
        dict with keys records/verts/idx/tverts/tidx holding int addresses."""
        info = self.info()
        records = np.zeros(info.block_count, RECORD_DTYPE)
        if into is None:
            verts = np.zeros(info.vertex_span, VERTEX_DTYPE); idx = np.zeros(info.index_span, np.uint32)
            tverts = np.zeros(info.trans_vertex_span, VERTEX_DTYPE); tidx = np.zeros(info.trans_index_span, np.uint32)
            self._check(self.L.vxb_result_download(self.h, _ptr(records), _ptr(verts), _ptr(idx), _ptr(tverts), _ptr(tidx)),
                        "vxb_result_download")
            return Result(self.n, info, records, verts, idx, tverts, tidx)
        self._check(self.L.vxb_result_download(self.h, _ptr(records), C.c_void_p(into["verts"]), C.c_void_p(into["idx"]),
                                               C.c_void_p(into["tverts"]), C.c_void_p(into["tidx"])), "vxb_result_download")
        return Result(self.n, info, records, None, None, None, None)
