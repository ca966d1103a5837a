#!/usr/bin/env python
"""bench.py - Mvoxels/s polygonized (BASELINE.json metric) on the seeded Perlin terrain.

  python bench.py --gpus N --steps K --warmup W            this repo's CUDA path (one rank per GPU under torchrun)
  python bench.py --impl reference --gpus N --steps K ...  the reference's own OpenMP CPU Polygonizer::Execute
                                                           (oracle/_ref, unmodified sources compiled by oracle/Makefile)

A "step" is one full polygonization (all LOD levels + transition cells = what the reference's Execute always
computes; BASELINE configs[2] asks for levels 0-3, a subset) of one n^3 grid.  Rank r works on its own tile of the
same endless terrain (x-origin shifted by r*n): independent objects, no data-path collective -> weak scaling.
  value : n^3 * ranks / time, grid resident in HBM, CUDA events on the launching stream, max over ranks.
  e2e   : same, through the C ABI with HOST (pinned) buffers: H2D of the 3 dense volumes and D2H of the full
          result (directory + vertex/index arenas) inside the timed region, every step.
Inputs (1 GiB per channel at 1024^3) are far larger than the 126 MB L2, so no explicit L2 flush is needed.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=1024, help="grid edge n (power of two)")
    ap.add_argument("--levels", type=int, default=0, help="LOD levels to compute (0 = all, as the reference)")
    ap.add_argument("--no-transitions", action="store_true")
    ap.add_argument("--cpu-sample-size", type=int, default=0, help="grid edge of the CPU-baseline sample (0 = same as --size, capped at 1024)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def measured_peak_hbm():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()  # the exact PID we started
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        clocks, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                clocks.append(float(r[0])); mx = float(r[1])
                for k, name in enumerate(names):
                    if r[2 + k].lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        clocks.sort()
        return {"sm_mhz": clocks[len(clocks) // 2] if clocks else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(clocks)}


def reference_run(n, steps, warmup, origin_x, torch, device_for_gen, budget_s=200.0):
    """Times the reference's own Polygonizer::Execute (all host threads) on the terrain tile; returns (Mvoxels/s, info).
    The sample is the full n^3 tile when warmup + steps executions fit the time budget, else the (n/2)^3 tile of the
    same terrain (bounded sample: the run must end within a few minutes whatever K is)."""
    import harness
    from voxels_b200 import synth
    if not os.path.exists(harness.REF_LIB):
        return None, {"unavailable": "oracle/_ref/libvxh_ref.so not built (make -C oracle ref needs the reference checkout)"}
    ref = harness.reference()
    # every host core this process may use (torchrun exports OMP_NUM_THREADS=1, which is not what the reference would run with)
    threads = max(ref.L.vxh_max_threads(), len(os.sched_getaffinity(0)))
    while True:
        dist, mat, blend = synth.terrain(n, device_for_gen, origin=(origin_x, 0))
        dist, mat, blend = dist.cpu().numpy(), mat.cpu().numpy(), blend.cpu().numpy()
        grid = ref.grid_from_dense(dist, mat, blend)
        times = []
        shrink = False
        for i in range(warmup + steps):
            s, sec = ref.polygonize(grid, threads=threads)
            ref.surface_destroy(s)
            if i == 0 and n > 256 and sec * (warmup + steps) > budget_s:
                shrink = True
                break
            if i >= warmup:
                times.append(sec)
        ref.grid_destroy(grid)
        if not shrink:
            break
        n //= 2
    per_step = sum(times) / len(times)
    return n ** 3 / per_step / 1e6, {"cores": threads, "seconds_per_execute": per_step, "best_seconds": min(times), "n": n}


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n = args.size

    import torch

    # ---------------------------------------------------------------- reference arm (CPU, rank 0 only)
    if args.impl == "reference":
        if rank != 0:
            return 0
        gen_dev = "cuda:0" if torch.cuda.is_available() else "cpu"
        sample_n = min(n, 1024)
        t0 = time.time()
        value, info = reference_run(sample_n, args.steps, args.warmup, 0, torch, gen_dev)
        if value is None:
            print(json.dumps({"impl": "reference", "unavailable": info["unavailable"]}))
            return 0
        line = {
            "impl": "reference", "metric": "Mvoxels/s polygonized", "value": value, "unit": "Mvoxels/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": info["seconds_per_execute"] * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int8 samples / fp32 vertices", "data": "synthetic",
            "config": {"workload": "%d^3 seeded Perlin terrain, all LOD levels + transition cells (reference Polygonizer::Execute)" % info["n"],
                       "impl": "unmodified reference sources, g++ -O2 -fopenmp -msse2, OMP threads = %d" % info["cores"]},
            "cpu_baseline": {"value": value, "unit": "Mvoxels/s", "cores": info["cores"], "kind": "reference",
                             "sample": "%s %d^3 grid, Polygonizer::Execute only (grid build excluded), mean of %d runs"
                                       % ("full" if info["n"] == sample_n else "bounded sample (time budget):", info["n"], args.steps)},
            "e2e": {"value": value, "unit": "Mvoxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.time() - t0,
        }
        print(json.dumps(line))
        return 0

    # ---------------------------------------------------------------- this repo's CUDA path
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: bench.py measures the sm_100a kernels and has no CPU fallback"}))
        return 2
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import voxels_b200
    from voxels_b200 import synth
    from voxels_b200.dist import Ranks, tile_origin, whole_job_throughput
    ranks = Ranks("nccl", dev)

    flags = voxels_b200.FLAG_NO_TRANSITIONS if args.no_transitions else 0
    dist, mat, blend = synth.terrain(n, dev, origin=tile_origin(rank, n))
    torch.cuda.synchronize()
    ctx = voxels_b200.Context(local_rank)
    ctx.set_device_grid(n, dist.data_ptr(), mat.data_ptr(), blend.data_ptr(), keep=(dist, mat, blend))
    stream = torch.cuda.ExternalStream(ctx.L.vxb_stream(ctx.h), device=dev)

    def barrier():
        ranks.barrier()
        torch.cuda.synchronize()

    max_over_ranks = ranks.max_over_ranks

    def timed(fn, steps):
        """K steps bracketed by barrier+synchronize; CUDA events on the launching stream; max over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)) / steps

    # -- value: grid resident in HBM --
    info = None

    def step_resident():
        nonlocal info
        info = ctx.polygonize(args.levels, flags)

    # clocks are sampled (nvidia-smi, 50 ms period) from the warm-up through the timed region; the timed region of a
    # few ms-long steps is shorter than one sample period, so identical untimed steps keep the load up for >= 1.5 s
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t_load = time.time()
    for _ in range(max(args.warmup, 3)):
        step_resident()
    ms_resident = timed(step_resident, args.steps)
    while rank == 0 and time.time() - t_load < 1.5:
        step_resident()
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["window"] = "warm-up + timed region + identical untimed steps, %.1f s under load" % (time.time() - t_load)
    launches_per_step = info.kernel_launches
    device_ms_inner = info.device_ms

    # -- per-kernel times for the roofline (separate steps: the extra events add a little stream overhead) --
    kind_ms = [0.0] * 8
    kind_launches = [0] * 8
    ksteps = max(3, min(args.steps, 10))
    for _ in range(ksteps):
        ctx.polygonize(args.levels, flags | voxels_b200.FLAG_KERNEL_TIMES)
        for k in range(8):
            ms, ln = ctx.kernel_ms(k)
            kind_ms[k] += ms / ksteps
            kind_launches[k] = ln
    info = ctx.polygonize(args.levels, flags)
    levels_computed = info.levels_computed
    V, I, TV, TI = info.vertex_total, info.index_total, info.trans_vertex_total, info.trans_index_total
    d_level0 = float(n) ** 3
    d_upper = sum((n >> l) ** 3 for l in range(1, levels_computed))
    bytes_out = 48.0 * (V + TV) + 4.0 * (I + TI)
    bytes_mat = 4.0 * (V + TV)
    bytes_alg_total = d_level0 + d_upper + bytes_mat + bytes_out  # D + M + O
    peak, peak_src = measured_peak_hbm()
    kinds = ["vxb_scan_kernel (streams the level-0 distance volume once)", "vxb_block_info/vxb_select kernels",
             "vxb_classify_kernel (one launch per level: tiles, case codes, material votes)",
             "vxb_decide_kernel (per block: ordering, reuse decisions, scans; all levels)",
             "vxb_vertex_kernel (flat, one thread per new vertex)", "vxb_triangle_kernel (flat, one thread per non-trivial cell)",
             "vxb_transition_kernel (per mid-level block)", "vxb_finish_kernel (compaction + directory)"]
    # SURVEY.md 8(d) split per kernel: level-0 samples are read once by the scan; coarser levels' samples by classify;
    # materials + vertices by the vertex kernel; indices by the triangle kernel; transition output by its kernel.
    # Tile re-reads of candidate blocks and the intermediate cell records are overhead, not algorithmic bytes.
    alg_by_kind = [d_level0, 0.0, float(d_upper), 0.0, 52.0 * V, 4.0 * I, 52.0 * TV + 4.0 * TI, 0.0]
    dom = max(range(8), key=lambda k: kind_ms[k])
    ach = alg_by_kind[dom] / (kind_ms[dom] * 1e-3) / 1e9 if kind_ms[dom] > 0 else 0.0
    # DRAM traffic of that kernel kind per step from the committed `ncu --set full` capture (only valid for the default workload)
    traffic, traffic_src = None, None
    try:
        kind_keys = ["scan", "select", "classify", "decide", "vertex", "triangle", "transition", "finish"]
        with open(os.path.join(REPO, "profiles", "r01d_traffic.json")) as f:
            tj = json.load(f)
        if n == 1024 and args.levels == 0 and not args.no_transitions:
            e = tj["per_kind"][kind_keys[dom]]
            traffic, traffic_src = e["dram_read"] + e["dram_write"], tj["capture"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": kinds[dom], "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": peak_src, "algorithmic_bytes_per_step": alg_by_kind[dom], "kernel_ms_per_step": kind_ms[dom],
                "launches_per_step": kind_launches[dom],
                "all_kernels": {kinds[k]: {"ms_per_step": kind_ms[k], "launches": kind_launches[k], "algorithmic_bytes": alg_by_kind[k],
                                           "achieved_gbs": (alg_by_kind[k] / (kind_ms[k] * 1e-3) / 1e9 if kind_ms[k] > 0 else 0.0)} for k in range(8)},
                "job": {"algorithmic_bytes": bytes_alg_total, "device_ms": device_ms_inner,
                        "achieved_gbs": bytes_alg_total / (device_ms_inner * 1e-3) / 1e9, "frac": bytes_alg_total / (device_ms_inner * 1e-3) / 1e9 / peak}}

    # -- e2e: host buffers through the C ABI, H2D + D2H inside the timed region --
    # headline e2e: the grid arrives in the reference's own storage form (the bytes of Grid::PackForSave: RLE blocks),
    # is copied as is and decoded on the GPU; e2e_dense: the same with three dense n^3 host volumes.
    e2e = None
    e2e_dense = None
    if not args.no_e2e:
        h_dist, h_mat, h_blend = (t.cpu() for t in (dist, mat, blend))
        packed = voxels_b200.pack_dense(h_dist.numpy(), h_mat.numpy(), h_blend.numpy())
        h_blob = torch.from_numpy(packed.copy()).pin_memory()
        blob_bytes = int(h_blob.numel())
        h_dist, h_mat, h_blend = (t.pin_memory() for t in (h_dist, h_mat, h_blend))
        ctx2 = voxels_b200.Context(local_rank)
        stream2 = torch.cuda.ExternalStream(ctx2.L.vxb_stream(ctx2.h), device=dev)
        out = {
            "verts": torch.empty(int(info.vertex_span * 1.1) * 48 + 4096, dtype=torch.uint8).pin_memory(),
            "idx": torch.empty(int(info.index_span * 1.1) * 4 + 4096, dtype=torch.uint8).pin_memory(),
            "tverts": torch.empty(int(info.trans_vertex_span * 1.1) * 48 + 4096, dtype=torch.uint8).pin_memory(),
            "tidx": torch.empty(int(info.trans_index_span * 1.1) * 4 + 4096, dtype=torch.uint8).pin_memory(),
        }
        into = {k: v.data_ptr() for k, v in out.items()}
        d2h = [0]

        parts = {"upload": 0.0, "polygonize": 0.0, "download": 0.0, "steps": 0}  # host clock around the three (synchronous) calls

        def finish_step(t0):
            t1 = time.perf_counter()
            i2 = ctx2.polygonize(args.levels, flags)
            t2 = time.perf_counter()
            ctx2.download(into=into)
            t3 = time.perf_counter()
            parts["upload"] += t1 - t0; parts["polygonize"] += t2 - t1; parts["download"] += t3 - t2; parts["steps"] += 1
            d2h[0] = i2.block_count * 128 + i2.vertex_span * 48 + i2.index_span * 4 + i2.trans_vertex_span * 48 + i2.trans_index_span * 4

        def step_e2e_packed():
            t0 = time.perf_counter()
            ctx2.upload_packed(h_blob.data_ptr(), blob_bytes)
            finish_step(t0)

        def step_e2e_dense():
            t0 = time.perf_counter()
            ctx2.upload_dense_ptr(n, h_dist.data_ptr(), h_mat.data_ptr(), h_blend.data_ptr())
            finish_step(t0)

        def breakdown():
            k = max(parts["steps"], 1)
            out = {name: round(1e3 * parts[name] / k, 3) for name in ("upload", "polygonize", "download")}
            parts.update(upload=0.0, polygonize=0.0, download=0.0, steps=0)
            return out

        esteps = max(2, min(args.steps, 5))
        stream_saved = stream
        stream = stream2
        for _ in range(2):
            step_e2e_packed()
        breakdown()
        ms_e2e = timed(step_e2e_packed, esteps)
        e2e = {"value": whole_job_throughput(n, world, ms_e2e), "unit": "Mvoxels/s", "ms_per_step": ms_e2e, "steps": esteps, "host_ms": breakdown(),
               "h2d_bytes_per_step": blob_bytes + 8 * (n // 16) ** 3, "d2h_bytes_per_step": int(d2h[0]),
               "path": "vxb_grid_upload_packed (PackForSave bytes, pinned host -> HBM, RLE decode on the GPU) + vxb_polygonize + "
                       "vxb_result_download (directory + arenas, HBM -> pinned host)"}
        for _ in range(2):
            step_e2e_dense()
        breakdown()
        ms_dense = timed(step_e2e_dense, esteps)
        e2e_dense = {"value": whole_job_throughput(n, world, ms_dense), "unit": "Mvoxels/s", "ms_per_step": ms_dense, "steps": esteps, "host_ms": breakdown(),
                     "h2d_bytes_per_step": 3 * n ** 3, "d2h_bytes_per_step": int(d2h[0]),
                     "path": "vxb_grid_upload_dense (3 dense volumes, pinned host -> HBM) + vxb_polygonize + vxb_result_download"}
        stream = stream_saved
        # the same calls from two host threads, each with its own context and buffers: step k's download (D2H) overlaps
        # step k+1's upload (H2D) and kernels, the way a client streams many grids through one GPU.  Extra information,
        # timed on the host clock over all steps; the headline `e2e` above is one step at a time.
        if True:
            ctx3 = voxels_b200.Context(local_rank)
            out3 = {k: torch.empty_like(v).pin_memory() for k, v in out.items()}
            into3 = {k: v.data_ptr() for k, v in out3.items()}
            psteps = max(4, esteps * 2)
            gate = threading.Barrier(3, timeout=300)
            failures = []

            def worker(c, buffers):
                try:
                    for i in range(psteps + 1):
                        if i == 1:
                            gate.wait()  # first step = warm-up
                        c.upload_packed(h_blob.data_ptr(), blob_bytes)
                        c.polygonize(args.levels, flags)
                        c.download(into=buffers)
                    gate.wait()
                except Exception as exc:  # never leave the other parties waiting
                    failures.append(repr(exc))
                    gate.abort()

            threads = [threading.Thread(target=worker, args=(ctx2, into)), threading.Thread(target=worker, args=(ctx3, into3))]
            for t in threads:
                t.start()
            try:
                gate.wait()
                t0 = time.perf_counter()
                gate.wait()
                wall_ms = 1e3 * (time.perf_counter() - t0)
            except threading.BrokenBarrierError:
                wall_ms = None
            for t in threads:
                t.join()
            if wall_ms is not None and not failures:
                ms_pipe = max_over_ranks(wall_ms / (2 * psteps))
                e2e["two_in_flight"] = {"value": whole_job_throughput(n, world, ms_pipe), "unit": "Mvoxels/s", "ms_per_step": ms_pipe, "steps": 2 * psteps,
                                        "timing": "host clock over all steps; two host threads, one context each, same calls as e2e"}
            else:
                max_over_ranks(0.0)  # keep the ranks' collectives aligned
                e2e["two_in_flight"] = {"value": None, "error": "; ".join(failures) or "barrier timeout"}
            ctx3.close()
            del out3
        ctx2.close()
        del h_dist, h_mat, h_blend, h_blob, out

    # -- CPU baseline: the reference itself on this box's host cores (rank 0, N=1 only) --
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample_n = args.cpu_sample_size or min(n, 1024)
        t0 = time.time()
        v, ci = reference_run(sample_n, 2, 0, 0, torch, dev)
        if v is not None:
            cpu = {"value": v, "unit": "Mvoxels/s", "cores": ci["cores"], "kind": "reference",
                   "sample": "%s %d^3 terrain tile (same bytes as the GPU step when sizes match), Polygonizer::Execute only, mean of 2 runs, %.1f s wall incl. grid build"
                             % ("full" if ci["n"] == sample_n else "bounded sample:", ci["n"], time.time() - t0)}
        else:
            cpu = {"value": None, "unit": "Mvoxels/s", "cores": 0, "kind": "reference", "sample": ci["unavailable"]}

    value = whole_job_throughput(n, world, ms_resident)
    if rank == 0:
        line = {
            "metric": "Mvoxels/s polygonized", "value": value, "unit": "Mvoxels/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_resident, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int8 samples / fp32 vertices", "data": "synthetic",
            "config": {"workload": "%d^3 seeded Perlin terrain per GPU, LOD levels 0-%d%s (BASELINE configs[2] asks for levels 0-3 + transitions)"
                                   % (n, levels_computed - 1, "" if args.no_transitions else " + transition cells"),
                       "grid": "dense int8 distance + uint8 material + uint8 blend, resident in HBM", "sharding": "one independent terrain tile per rank, no data-path collective",
                       "l2": "inputs (%.2f GiB per channel) larger than the 126 MB L2; no flush" % (n ** 3 / 2.0 ** 30),
                       "vertices": int(V), "indices": int(I), "transition_vertices": int(TV), "transition_indices": int(TI), "blocks_emitted": int(info.block_count)},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "e2e_dense": e2e_dense, "clocks": clocks,
            "gpu_launches": int(launches_per_step * args.steps), "device_ms_per_step_inner": device_ms_inner,
        }
        print(json.dumps(line))
    ctx.close()
    ranks.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
