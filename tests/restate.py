"""ctypes driver for oracle/restate/vxb_restate.cpp (CPU restatement, test infrastructure)."""
import ctypes as C
import os

import numpy as np

from harness import LevelDump, ROW_DTYPE, VERTEX_DTYPE, REPO

RESTATE_LIB = os.path.join(REPO, "build", "oracle", "libvxr_restate.so")


class Restate:
    def __init__(self, path=RESTATE_LIB):
        L = self.L = C.CDLL(path)
        vp, u = C.c_void_p, C.c_uint
        L.vxr_run.restype = vp; L.vxr_run.argtypes = [u, vp, vp, vp, vp, vp, C.c_int]
        L.vxr_destroy.restype = None; L.vxr_destroy.argtypes = [vp]
        L.vxr_levels.restype = u; L.vxr_levels.argtypes = [vp]
        L.vxr_blocks.restype = u; L.vxr_blocks.argtypes = [vp, u]
        L.vxr_stats.restype = None; L.vxr_stats.argtypes = [vp, vp]
        L.vxr_empty_flags.restype = None; L.vxr_empty_flags.argtypes = [vp, vp]
        L.vxr_level_totals.restype = None; L.vxr_level_totals.argtypes = [vp, u, vp]
        L.vxr_level_dump.restype = None; L.vxr_level_dump.argtypes = [vp, u, vp, vp, vp, vp, vp]

    @staticmethod
    def _ptr(a):
        return None if a is None else a.ctypes.data_as(C.c_void_p)

    def run(self, dist, mat, blend, material_table=None, valid_mask=None, max_levels=0):
        n = dist.shape[0]
        self._keep = (dist, mat, blend)
        return self.L.vxr_run(n, self._ptr(dist), self._ptr(mat), self._ptr(blend), self._ptr(material_table),
                              self._ptr(valid_mask), max_levels)

    def levels(self, h):
        return self.L.vxr_levels(h)

    def stats(self, h):
        st = np.zeros(20, np.uint32)
        self.L.vxr_stats(h, self._ptr(st))
        return st

    def empty_flags(self, h, n):
        f = np.zeros((n // 16) ** 3, np.uint8)
        self.L.vxr_empty_flags(h, self._ptr(f))
        return f

    def level(self, h, level):
        nb = self.L.vxr_blocks(h, level)
        totals = np.zeros(4, np.uint64)
        self.L.vxr_level_totals(h, level, self._ptr(totals))
        rows = np.zeros(nb, ROW_DTYPE)
        verts = np.zeros(int(totals[0]), VERTEX_DTYPE); idx = np.zeros(int(totals[1]), np.uint32)
        tverts = np.zeros(int(totals[2]), VERTEX_DTYPE); tidx = np.zeros(int(totals[3]), np.uint32)
        self.L.vxr_level_dump(h, level, self._ptr(rows), self._ptr(verts), self._ptr(idx), self._ptr(tverts), self._ptr(tidx))
        return LevelDump(rows, verts, idx, tverts, tidx)

    def destroy(self, h):
        self.L.vxr_destroy(h)
