"""voxels_b200 - B200-native (sm_100a) Transvoxel polygonizer backend for stoyannk/voxels.

The product is native: CUDA kernels + a C ABI (include/vxb200.h -> voxels_b200/lib/libvxb200.so) and a
C++ drop-in for the reference's Polygonizer API (libvoxels_b200.so).  This package is the thin Python
host side (ctypes) used by the tests and bench.py.  There is no CPU fallback: importing works without
a GPU, creating a `Context` does not.
"""
from .capi import (Context, Result, VxbError, library_path, load_library, pack_dense, FLAG_NO_TRANSITIONS, FLAG_KERNEL_TIMES,
                   RECORD_DTYPE, VERTEX_DTYPE)

__all__ = ["Context", "Result", "VxbError", "library_path", "load_library", "pack_dense", "FLAG_NO_TRANSITIONS",
           "FLAG_KERNEL_TIMES", "RECORD_DTYPE", "VERTEX_DTYPE"]
