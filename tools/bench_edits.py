#!/usr/bin/env python
"""BASELINE configs[4]: n^3 terrain + K seeded sphere add/subtract edits, incremental re-polygonization after each.

Both sides run the SAME client code (tests/harness/vxh_capi.cpp over the reference's public API):
    Grid::InjectSurface(...)  ->  Modification{Map, Min/MaxCornerModified}  ->  Polygonizer::Execute(grid, materials, modification)
once linked with the unmodified reference (oracle/_ref/libvxh_ref.so, all host threads) and once with this repo's
drop-in backend (build/libvxh_b200.so).  Timed per edit: Execute only (the InjectSurface call is the reference's own
host code on both sides and is reported separately).  The drop-in's Execute includes reading the touched blocks from the
Grid, the H2D update, all kernels, the D2H of the re-created blocks and the splice into the surface.

A third column, "device": the same edits WITHOUT the host grid - vxb_grid_inject_surface (the edit kernel on the
device-resident grid) + vxb_polygonize_region + vxb_result_download through the C ABI; timed per edit as
one unit (edit + re-polygonization + the re-created blocks back on the host) and compared with the reference's
re-created blocks at every check point.

    python tools/bench_edits.py [--size 512] [--edits 1000] [--check-every 100]
"""
import argparse
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

import numpy as np  # noqa: E402


def edit_sequence(n, count, dist, seed=42):
    """LCG-free equivalent of SURVEY.md 8(d) config 5: integer centres near the surface, r in {4,6,8,10}, alternating add/subtract."""
    rng = np.random.RandomState(seed)
    edits = []
    while len(edits) < count:
        x, y, z = (int(v) for v in rng.randint(32, n - 32, size=3))
        if abs(int(dist[z, y, x])) >= 4:
            continue  # resample until the centre is near the surface
        r = int(rng.choice([4, 6, 8, 10]))
        edits.append(((x, y, z), float(r), float(2 * r + 4), 0 if len(edits) % 2 == 0 else 2))
    return edits


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--edits", type=int, default=1000)
    ap.add_argument("--check-every", type=int, default=100, help="compare the two surfaces every N edits (0 = never)")
    args = ap.parse_args()
    import torch
    import compare
    import harness
    from voxels_b200 import synth
    n = args.size
    dev = "cuda:0" if torch.cuda.is_available() else "cpu"
    dist, mat, blend = (t.cpu().numpy() for t in synth.terrain(n, dev))
    edits = edit_sequence(n, args.edits, dist)
    libs = [("reference", harness.reference()), ("b200", harness.load(harness.B200_LIB))]
    state = {}
    # the device column: grid resident in HBM, edited there
    import voxels_b200
    from voxels_b200 import capi
    dctx = voxels_b200.Context(0)
    dctx.upload_dense(dist, mat, blend)
    dctx.polygonize()
    dev = {"seconds": 0.0, "inject": 0.0, "polygonize": 0.0, "download": 0.0, "blocks": 0, "mismatches": 0}
    # page-locked result buffers for the re-created blocks (a few hundred KB per edit; sized generously once)
    pin = {"verts": torch.empty(64 << 20, dtype=torch.uint8).pin_memory(), "idx": torch.empty(32 << 20, dtype=torch.uint8).pin_memory(),
           "tverts": torch.empty(32 << 20, dtype=torch.uint8).pin_memory(), "tidx": torch.empty(16 << 20, dtype=torch.uint8).pin_memory()}
    into = {k: v.data_ptr() for k, v in pin.items()}
    for name, lib in libs:
        g = lib.grid_from_dense(dist, mat, blend)
        t0 = time.time()
        s, sec = lib.polygonize(g)
        state[name] = {"lib": lib, "grid": g, "surface": s, "mod": lib.modification_create(), "full_s": sec, "exec_s": 0.0, "inject_s": 0.0,
                       "blocks": 0}
    mismatches = 0
    for i, (pos, radius, extent, kind) in enumerate(edits):
        for name, lib in libs:
            st = state[name]
            t0 = time.perf_counter()
            box = lib.grid_inject_sphere(st["grid"], pos, radius, extent, kind)
            t1 = time.perf_counter()
            before = len(lib.modification_blocks(st["mod"])) if i % 100 == 0 else None
            s2, sec = lib.polygonize(st["grid"], modification=st["mod"], surface=st["surface"], box=box)
            st["inject_s"] += t1 - t0
            st["exec_s"] += sec
        sphere = capi.Surface.sphere((0, 0, 0), radius)
        p32, e32 = np.array(pos, np.float32), np.full(3, extent, np.float32)
        t0 = time.perf_counter()
        dbox = dctx.inject_surface(p32, e32, sphere, kind)
        t1 = time.perf_counter()
        dctx.polygonize_region(dbox[:3], dbox[3:])
        t2 = time.perf_counter()
        part = dctx.download(into=into)
        t3 = time.perf_counter()
        dev["seconds"] += t3 - t0; dev["inject"] += t1 - t0; dev["polygonize"] += t2 - t1; dev["download"] += t3 - t2
        dev["blocks"] += len(part.records)
        if args.check_every and (i + 1) % args.check_every == 0:
            part = dctx.download()   # numpy copies for the comparison (untimed)
            # the blocks this edit re-created = the tail of every level of the reference's surface (erase + append)
            a = state["reference"]
            for l in range(a["lib"].surface_levels(a["surface"])):
                want, got = a["lib"].surface_level(a["surface"], l), part.level(l)
                k = len(got.rows); cut = len(want.rows) - k
                r = want.rows
                tail = harness.LevelDump(r[cut:], want.verts[int(r["nv"][:cut].sum()):], want.idx[int(r["ni"][:cut].sum()):],
                                         want.tverts[int(r["tnv"][:cut].sum()):], want.tidx[int(r["tni"][:cut].sum()):])
                if cut < 0 or compare.level_diff(tail, got, "device edit %d L%d" % (i, l)):
                    dev["mismatches"] += 1
        if args.check_every and (i + 1) % args.check_every == 0:
            a, b = state["reference"], state["b200"]
            for l in range(a["lib"].surface_levels(a["surface"])):
                if compare.level_diff(a["lib"].surface_level(a["surface"], l), b["lib"].surface_level(b["surface"], l), "edit %d L%d" % (i, l)):
                    mismatches += 1
    out = {"metric": "edits/s (incremental re-polygonization, Execute only)", "config": {"workload": "%d^3 seeded Perlin terrain, %d sphere add/subtract edits" % (n, len(edits))},
           "parity_checks_failed": mismatches}
    for name, lib in libs:
        st = state[name]
        blocks = len(lib.modification_blocks(st["mod"]))
        out[name] = {"edits_per_s": len(edits) / st["exec_s"], "ms_per_edit": 1e3 * st["exec_s"] / len(edits), "full_polygonize_s": st["full_s"],
                     "inject_ms_per_edit": 1e3 * st["inject_s"] / len(edits), "blocks_recreated_per_edit": blocks / len(edits),
                     "threads": lib.L.vxh_max_threads() if name == "reference" else None}
    out["speedup_execute"] = out["b200"]["edits_per_s"] / out["reference"]["edits_per_s"]
    out["device"] = {"edits_per_s": len(edits) / dev["seconds"], "ms_per_edit": 1e3 * dev["seconds"] / len(edits), "blocks_recreated_per_edit": dev["blocks"] / len(edits),
                     "parity_checks_failed": dev["mismatches"],
                     "host_ms": {k: round(1e3 * dev[k] / len(edits), 3) for k in ("inject", "polygonize", "download")},
                     "what": "vxb_grid_inject_surface (edit kernel on the device grid) + vxb_polygonize_region + vxb_result_download, host clock per edit; "
                             "compare with reference inject_ms_per_edit + ms_per_edit"}
    out["speedup_device_vs_reference_inject_plus_execute"] = (out["reference"]["ms_per_edit"] + out["reference"]["inject_ms_per_edit"]) / out["device"]["ms_per_edit"]
    dctx.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
