#!/usr/bin/env python
"""BASELINE configs[3]: ONE n^3 terrain polygonized by all ranks (z-slabs, one GPU each).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/bench_sharded.py [--size 2048] [--steps 10] [--warmup 3] [--verify]

Every rank fills its own slab of the cube (synthetic terrain, on the device), maps the peers' slabs over NVLink and runs
voxels_b200.dist.ShardedGrid.polygonize: scan + nested levels, ONE in-place NCCL all-gather of the last nested level's
material pages, the rest.  Timed per step on the device (CUDA events on the context's stream), max over ranks; the block
directories are all-gathered after the timed loop.  --verify (needs the whole grid to fit rank 0's GPU next to its slab)
re-runs the grid unsharded on rank 0 and compares the gathered directory entry by entry and rank 0's geometry bit by bit.
Rank 0 prints one JSON line."""
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
    ap.add_argument("--verify", action="store_true")
    ap.add_argument("--balance", type=int, default=0, metavar="PLANES",
                    help="after a first run with equal slabs, re-shard with boundaries (multiples of PLANES) that even out the work it measured")
    args = ap.parse_args()
    import numpy as np
    import torch
    import voxels_b200
    from voxels_b200 import capi, synth
    from voxels_b200.dist import Ranks, ShardedGrid, slab_planes, balanced_planes, layer_weights_from_directory

    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    ranks = Ranks("nccl", dev)
    n = args.size
    def build(planes):
        g = ShardedGrid(ranks, n, planes=planes)
        z0, z1 = slab_planes(n, ranks.rank, ranks.world, planes)
        synth.terrain(n, dev, z_range=(z0, z1), out=g.slab_tensors())
        g.ready()
        return g

    sg = build(None)
    planes = None
    if args.balance and ranks.world > 1:
        # z is up: a terrain's surface sits in a few z-layers.  One run with equal slabs tells where the work is; the
        # slabs are then re-cut so that every rank gets the same share (what a client does from the previous frame).
        sg.polygonize()
        directory, _ = sg.directory()
        planes = balanced_planes(layer_weights_from_directory(n, directory), ranks.world, args.balance // 16)
        sg.close()
        torch.cuda.synchronize(dev)
        ranks.barrier()
        sg = build(planes)

    stream = torch.cuda.ExternalStream(sg.ctx.stream(), device=dev)
    times, mine_ms, inner_ms = [], [], []
    for step in range(args.warmup + args.steps):
        ranks.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        info = sg.polygonize()
        e1.record(stream)
        torch.cuda.synchronize(dev)
        ms = ranks.max_over_ranks(e0.elapsed_time(e1))
        if step >= args.warmup:
            times.append(ms)
            mine_ms.append(e0.elapsed_time(e1))
            inner_ms.append(info.device_ms)
    directory, owner = sg.directory()
    ms = float(np.mean(times))
    per_rank = torch.zeros(ranks.world * 2, dtype=torch.float64, device=dev)
    if ranks.td is not None:
        ranks.td.all_gather_into_tensor(per_rank, torch.tensor([float(np.mean(mine_ms)), float(np.mean(inner_ms))], dtype=torch.float64, device=dev))
    else:
        per_rank = torch.tensor([float(np.mean(mine_ms)), float(np.mean(inner_ms))])
    per_rank = per_rank.cpu().numpy().reshape(-1, 2)
    out = {"metric": "Mvoxels/s polygonized (one grid sharded over the GPUs)", "value": float(n) ** 3 / (ms * 1e-3) / 1e6, "unit": "Mvoxels/s",
           "n_gpus": ranks.world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "scaling": "strong",
           "config": {"workload": "%d^3 seeded Perlin terrain, all LOD levels + transition cells, z-slabs of %d planes" % (n, n // ranks.world),
                      "slab_boundaries": planes if planes is not None else "equal",
                      "exchange": "one NCCL exchange (material pages of the last nested level) + directory all-gather"},
           "blocks_total": int(len(directory)), "blocks_per_rank": [int((owner == r).sum()) for r in range(ranks.world)],
           "ms_per_rank": [round(float(v), 3) for v in per_rank[:, 0]], "device_ms_per_rank": [round(float(v), 3) for v in per_rank[:, 1]],
           "vertices_this_rank0": int(info.vertex_total)}

    if args.verify:
        problems = []
        mine = sg.ctx.download()
        if ranks.rank == 0:
            import compare
            ctx = voxels_b200.Context(local)
            full = synth.terrain(n, dev)
            ctx.set_device_grid(n, full[0].data_ptr(), full[1].data_ptr(), full[2].data_ptr(), keep=full)
            ctx.polygonize()
            single = ctx.download()
            if len(single.records) != len(directory):
                problems.append("directory sizes differ: %d vs %d" % (len(single.records), len(directory)))
            else:
                for f in ("level", "coord_id", "id", "vertex_count", "index_count", "trans_vertex_count", "trans_index_count"):
                    if not np.array_equal(single.records[f], directory[f]):
                        problems.append("directory field %s differs" % f)
            # geometry of rank 0's own blocks, bit by bit
            keys = set(zip(mine.records["level"].tolist(), mine.records["coord_id"].tolist()))
            keep = np.array([(l, c) in keys for l, c in zip(single.records["level"].tolist(), single.records["coord_id"].tolist())])
            sub = capi.Result.__new__(capi.Result)
            sub.n, sub.info, sub.records = n, None, single.records[keep]
            sub.verts, sub.idx, sub.tverts, sub.tidx, sub.stats = single.verts, single.idx, single.tverts, single.tidx, single.stats
            for l in range(info.levels_total):
                problems += compare.level_diff(sub.level(l), mine.level(l), "rank0 L%d" % l)
            ctx.close()
        out["verify_problems"] = problems[:10]
    if ranks.rank == 0:
        print(json.dumps(out))
    sg.close()
    ranks.close()


if __name__ == "__main__":
    main()
