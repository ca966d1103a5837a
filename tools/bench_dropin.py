#!/usr/bin/env python
"""Polygonizer::Execute of the drop-in (libvoxels_b200.so through tests/harness) on a reference Grid, host in / host out,
repeated, with the host-side stage clock of every execution (voxels_b200_last_execute_stages).

    python tools/bench_dropin.py [--size 1024] [--runs 10]

Prints one JSON line: per-run wall ms + stages, so that outliers (first-touch, thread wake-ups, NUMA) can be told apart
from the steady state.  The grid bytes come from the device fill (vxb_grid_fill), as in bench.py."""
import argparse
import ctypes
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--runs", type=int, default=10)
    args = ap.parse_args()
    import voxels_b200
    from voxels_b200 import capi
    import harness
    n = args.size
    ctx = voxels_b200.Context(0)
    ctx.fill(n, capi.Surface.terrain(n))
    hd, hm, hb = ctx.download_dense()
    ctx.close()
    dl = harness.load(harness.B200_LIB)
    g = dl.grid_from_dense(hd, hm, hb)
    del hd, hm, hb
    lib = ctypes.CDLL(os.path.join(REPO, "voxels_b200", "lib", "libvoxels_b200.so"))
    names = ("materials", "block_offsets", "gather_blocks", "upload_decode_tail", "kernels", "download_views")
    runs = []
    surf = None
    for i in range(args.runs):
        if surf is not None:
            dl.surface_destroy(surf)
        surf, sec = dl.polygonize(g)
        st6 = (ctypes.c_double * 6)()
        lib.voxels_b200_last_execute_stages(st6)
        runs.append({"ms": round(1e3 * sec, 3), **{k: round(v, 3) for k, v in zip(names, st6)}})
    dl.surface_destroy(surf); dl.grid_destroy(g)
    ms = sorted(r["ms"] for r in runs[1:])
    print(json.dumps({"size": n, "omp": {k: os.environ.get(k) for k in ("OMP_PROC_BIND", "OMP_NUM_THREADS", "OMP_WAIT_POLICY", "OMP_PLACES")},
                      "cpus": len(os.sched_getaffinity(0)), "median_ms": ms[len(ms) // 2], "best_ms": ms[0], "worst_ms": ms[-1], "runs": runs}))


if __name__ == "__main__":
    main()
