#!/usr/bin/env python
"""BASELINE configs[3] and the strong-scaling line: ONE n^3 terrain polygonized by all ranks (voxels_b200.dist.ShardedGrid).

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_sharded.py --size 2048 [--verify] [--reference]

--verify     every rank also polygonizes the whole grid alone (single-GPU path) and compares ITS OWN blocks of the
             sharded run bit by bit (vertices, indices, transition meshes); rank 0 compares the merged directory entry
             by entry and the summed statistics.
--reference  (n <= 1024) rank 0 additionally runs the unmodified reference (oracle/_ref) on the same bytes and compares
             the single-GPU result level by level - with --verify that pins every rank's geometry to the reference.
Prints one JSON line on rank 0.
"""
import argparse
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=2048)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--group-planes", type=int, default=0)
    ap.add_argument("--mode", default="replicated", choices=["replicated", "cube"])
    ap.add_argument("--verify", action="store_true")
    ap.add_argument("--reference", action="store_true")
    args = ap.parse_args()
    import numpy as np
    import torch
    import voxels_b200
    from voxels_b200 import capi
    from voxels_b200.dist import Ranks, ShardedGrid

    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    ranks = Ranks("nccl", dev)
    n = args.size
    sg = ShardedGrid(ranks, n, group_planes=args.group_planes or None, mode=args.mode)
    terrain = capi.Surface.terrain(n)
    sg.fill(terrain)   # Grid::Create on the device: the whole grid on every rank (replicated) or the pieces it backs (cube)
    sg.ready()
    stream = torch.cuda.ExternalStream(sg.ctx.stream(), device=dev)
    for _ in range(args.warmup):
        info = sg.polygonize()
    inner = []
    ranks.barrier(); torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        assert sg.ctx.polygonize_sharded(3, 0) == 0
        inner.append(sg.ctx.info().device_ms)
    e1.record(stream)
    ranks.barrier(); torch.cuda.synchronize(dev)
    ms = ranks.max_over_ranks(e0.elapsed_time(e1)) / args.steps
    info = sg.ctx.info()
    directory, owner = sg.directory()
    per_rank = torch.zeros(ranks.world, dtype=torch.float64, device=dev)
    mine_ms = torch.tensor([float(np.mean(inner))], dtype=torch.float64, device=dev)
    if ranks.td is not None:
        ranks.td.all_gather_into_tensor(per_rank, mine_ms)
    else:
        per_rank = mine_ms
    stats = torch.tensor([int(v) for v in info.stats], dtype=torch.int64, device=dev)
    if ranks.td is not None:
        ranks.td.all_reduce(stats)
    stats = (stats.cpu().numpy() & 0xFFFFFFFF).astype(np.uint32)
    out = {"metric": "Mvoxels/s polygonized (one grid over the GPUs)", "value": float(n) ** 3 / (ms * 1e-3) / 1e6, "unit": "Mvoxels/s",
           "n_gpus": ranks.world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "scaling": "strong",
           "config": {"workload": "%d^3 seeded Perlin terrain, all LOD levels + transition cells, volumes %s, scan groups / pieces of %d planes dealt cyclically" % (n, args.mode, sg.group_planes),
                      "exchange": "ncclAllGather of the per-block info + peer stores of material pages (ordered by a device-side barrier over the mapped buffers), inside every step"},
           "blocks_total": int(len(directory)), "blocks_per_rank": [int((owner == r).sum()) for r in range(ranks.world)],
           "vertices_per_rank": [int(directory["vertex_count"][owner == r].sum()) for r in range(ranks.world)],
           "device_ms_per_rank": [round(float(v), 3) for v in per_rank.cpu().numpy()], "launches_per_step": int(info.kernel_launches)}

    # where the step goes, per rank: one extra step with per-kernel events (plain launches, one stream)
    kinds = ["scan", "coarse lattices + block walk (info, pyramid, plan, select)", "block kernels levels>=1 + coarse", "block kernel level 0", "vertices levels>=1",
             "triangles levels>=1", "transitions", "finish", "exchange 0 (lattice publish + all-gather)", "exchange 1 (page publish + peer barrier)"]
    ranks.barrier()
    sg.ctx.polygonize_sharded(3, voxels_b200.FLAG_KERNEL_TIMES)
    if True:
        mine_k = torch.tensor([sg.ctx.kernel_ms(k)[0] for k in range(10)], dtype=torch.float64, device=dev)
        all_k = torch.zeros(ranks.world * 10, dtype=torch.float64, device=dev)
        if ranks.td is not None:
            ranks.td.all_gather_into_tensor(all_k, mine_k)
        else:
            all_k = mine_k
        all_k = all_k.cpu().numpy().reshape(ranks.world, 10)
        out["kernel_ms_per_rank"] = {kinds[k]: [round(float(v), 3) for v in all_k[:, k]] for k in range(10)}
    if args.verify:
        import compare
        problems = []
        mine = sg.ctx.download()
        ctx = voxels_b200.Context(local)
        ctx.fill(n, terrain)
        sinfo = ctx.polygonize()
        single = ctx.download()
        # this rank's blocks of the sharded run against the same blocks of the single-GPU run, bit by bit
        keys = set(zip(mine.records["level"].tolist(), mine.records["coord_id"].tolist()))
        keep = np.array([(l, c) in keys for l, c in zip(single.records["level"].tolist(), single.records["coord_id"].tolist())], bool)
        sub = capi.Result.__new__(capi.Result)
        sub.n, sub.info, sub.records = n, None, single.records[keep]
        sub.verts, sub.idx, sub.tverts, sub.tidx, sub.stats = single.verts, single.idx, single.tverts, single.tidx, single.stats
        for l in range(info.levels_total):
            problems += compare.level_diff(sub.level(l), mine.level(l), "rank%d L%d" % (ranks.rank, l))
        if ranks.rank == 0:
            if len(single.records) != len(directory):
                problems.append("directory sizes differ: %d vs %d" % (len(single.records), len(directory)))
            else:
                for f in ("level", "coord_id", "id", "vertex_count", "index_count", "trans_vertex_count", "trans_index_count"):
                    if not np.array_equal(single.records[f], directory[f]):
                        problems.append("directory field %s differs" % f)
            if not np.array_equal(stats, single.stats):
                problems.append("summed statistics differ: %s vs %s" % (stats.tolist(), single.stats.tolist()))
            if args.reference and n <= 1024:
                import harness
                ref = harness.reference()
                hd, hm, hb = ctx.download_dense()
                g = ref.grid_from_dense(hd, hm, hb)
                s, sec = ref.polygonize(g, threads=max(ref.L.vxh_max_threads(), len(os.sched_getaffinity(0))))
                for l in range(ref.surface_levels(s)):
                    problems += compare.level_diff(ref.surface_level(s, l), single.level(l), "reference L%d" % l)
                if not np.array_equal(ref.surface_stats(s), single.stats):
                    problems.append("reference statistics differ")
                out["reference_seconds"] = sec
                ref.surface_destroy(s); ref.grid_destroy(g)
        bad = ranks.sum_over_ranks(len(problems))
        out["verify"] = {"ranks_compared": ranks.world, "problems_all_ranks": int(bad), "first_rank0": problems[:5],
                         "what": "every rank: its blocks vs the single-GPU run (bit-exact); rank 0: merged directory + summed statistics"
                                 + ("; single-GPU run vs the unmodified reference" if args.reference and n <= 1024 else "")}
        out["single_gpu_ms"] = float(sinfo.device_ms)
        ctx.close()
    if ranks.rank == 0:
        print(json.dumps(out))
    sg.close()
    ranks.close()


if __name__ == "__main__":
    main()
