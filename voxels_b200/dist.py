"""torch.distributed plumbing for the multi-GPU runs (one process per GPU; NCCL on the B200 box, gloo in the CPU tests).

Two ways to use N GPUs:
  * independent tiles (bench.py's weak-scaling line): every rank owns its own n^3 tile of the terrain (`tile_origin`);
    no data-path collective, only the barrier around the timed region and the max-reduction of the device times;
  * ONE grid sharded into z-slabs (`ShardedGrid`, SURVEY.md section 8e / BASELINE configs[3]): the cube lives in one
    virtual address range per rank whose slabs are the ranks' HBM, mapped into every peer over NVLink (file
    descriptors travel over Unix sockets, `exchange_fds`); the data path has ONE all-gather (the material pages of the
    last level that nests in a slab) and the directories are all-gathered at the end."""
import os
import socket
import threading


class Ranks:
    def __init__(self, backend=None, device=None):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.device = device
        self.td = None
        if self.world > 1:
            import torch.distributed as td
            self.td = td
            if not td.is_initialized():
                kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
                td.init_process_group(backend=backend or "nccl", **kw)

    def barrier(self):
        if self.td is not None:
            if self.device is not None and self.device.type == "cuda":
                self.td.barrier(device_ids=[self.local_rank])
            else:
                self.td.barrier()

    def max_over_ranks(self, value):
        """The slowest rank's value (multi-GPU times are reported as the max over ranks)."""
        if self.td is None:
            return float(value)
        import torch
        t = torch.tensor([float(value)], dtype=torch.float64, device=self.device if self.device is not None else "cpu")
        self.td.all_reduce(t, op=self.td.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, value):
        if self.td is None:
            return float(value)
        import torch
        t = torch.tensor([float(value)], dtype=torch.float64, device=self.device if self.device is not None else "cpu")
        self.td.all_reduce(t, op=self.td.ReduceOp.SUM)
        return float(t.item())

    def close(self):
        if self.td is not None and self.td.is_initialized():
            self.td.destroy_process_group()


def tile_origin(rank, n):
    """(x, y) origin of rank's tile in the endless terrain: tiles are laid out along x, n voxels apart."""
    return (rank * n, 0)


def whole_job_throughput(n, world, ms_per_step):
    """Mvoxels/s of the whole job: every rank polygonizes its own n^3 tile per step."""
    return float(n) ** 3 * world / (ms_per_step * 1e-3) / 1e6


# ---------------------------------------------------------------------------------------------------------------------
# one grid, z-slabs across the ranks
# ---------------------------------------------------------------------------------------------------------------------

def uniform_planes(n, world):
    return [n * r // world for r in range(world + 1)]


def slab_planes(n, rank, world, planes=None):
    """[z0, z1) of rank's slab."""
    planes = uniform_planes(n, world) if planes is None else planes
    return int(planes[rank]), int(planes[rank + 1])


def split_level(n, world, planes=None):
    """Levels [0, split) have blocks that nest in every slab (vxb_shard_exchange.split_level): the block edge 16 << l
    divides every slab boundary."""
    planes = uniform_planes(n, world) if planes is None else planes
    levels = (n // 16).bit_length()
    s = 0
    while s < levels and all(int(p) % (16 << s) == 0 for p in planes[1:-1]):
        s += 1
    return s


def balanced_planes(layer_weights, world, align_layers):
    """Slab boundaries that even out the work: layer_weights[i] = work in level-0 block layer i (16 planes), e.g. the
    vertices + triangles the last run produced there.  Boundaries are multiples of align_layers layers (that decides
    which levels still nest, see split_level); every slab gets at least one such unit.  Minimises the maximum slab
    weight (dynamic programme over the <= n/16/align_layers cells)."""
    import numpy as np
    w = np.asarray(layer_weights, np.float64)
    cells = len(w) // align_layers
    assert cells >= world, "more ranks than alignment units"
    cw = np.concatenate([[0.0], np.cumsum(w.reshape(cells, align_layers).sum(axis=1))])
    inf = float("inf")
    best = [[inf] * (cells + 1) for _ in range(world + 1)]
    cut = [[0] * (cells + 1) for _ in range(world + 1)]
    best[0][0] = 0.0
    for r in range(1, world + 1):
        for e in range(r, cells - (world - r) + 1):
            for b in range(r - 1, e):
                if best[r - 1][b] == inf:
                    continue
                # a tiny per-cell cost keeps empty regions spread over the ranks instead of piled on one
                cost = max(best[r - 1][b], cw[e] - cw[b] + 1e-9 * (e - b))
                if cost < best[r][e]:
                    best[r][e], cut[r][e] = cost, b
    bounds, e = [cells], cells
    for r in range(world, 0, -1):
        e = cut[r][e]
        bounds.append(e)
    return [b * align_layers * 16 for b in reversed(bounds)]


def layer_weights_from_directory(n, records, scan_units_per_byte=1.0 / 512):
    """Work per level-0 block layer from a block directory (numpy RECORD_DTYPE): a block's vertices + indices/3, spread
    over the layers it covers, plus the volume term - every layer is streamed once by the scan kernel whether or not
    the surface crosses it (measured on a B200: ~0.1 ns per vertex/triangle unit, scan at ~5.3 TB/s => ~1/512 unit per
    byte of the layer's 16 n^2 distance samples)."""
    import numpy as np
    nb0 = n // 16
    w = np.full(nb0, scan_units_per_byte * 16.0 * n * n, np.float64)
    level = records["level"].astype(np.int64)
    nbl = nb0 >> level
    z = records["coord_id"].astype(np.int64) // (nbl * nbl)
    work = (records["vertex_count"].astype(np.float64) + records["index_count"] / 3.0
            + records["trans_vertex_count"].sum(axis=1) + records["trans_index_count"].sum(axis=1) / 3.0)
    for l in np.unique(level):
        sel = level == l
        span = 1 << int(l)
        per_layer = np.bincount(z[sel], weights=work[sel] / span, minlength=nb0 >> int(l))
        w += np.repeat(per_layer, span)
    return w


def exchange_fds(rank, world, fds, key):
    """Every rank offers `fds` (a list of open file descriptors) to every peer; returns {peer: [fds...]} with this
    process's duplicates of the peers' descriptors.  Unix sockets in the abstract namespace + SCM_RIGHTS; all ranks run
    on one host (one process per GPU of a box)."""
    if world == 1:
        return {}
    name = lambda r: "\0vxb200-%s-%d" % (key, r)  # noqa: E731
    server = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    server.bind(name(rank))
    server.listen(world)

    def serve():
        for _ in range(world - 1):
            conn, _addr = server.accept()
            with conn:
                conn.recv(1)
                socket.send_fds(conn, [b"f"], list(fds))
    t = threading.Thread(target=serve, daemon=True)
    t.start()
    got = {}
    for peer in range(world):
        if peer == rank:
            continue
        c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        for attempt in range(600):  # the peer may not be listening yet
            try:
                c.connect(name(peer))
                break
            except (ConnectionRefusedError, FileNotFoundError):
                import time
                time.sleep(0.05)
        else:
            raise RuntimeError("exchange_fds: rank %d never reached rank %d" % (rank, peer))
        with c:
            c.send(b"r")
            _msg, received, _flags, _addr = socket.recv_fds(c, 1, len(fds))
        if len(received) != len(fds):
            raise RuntimeError("exchange_fds: expected %d descriptors from rank %d, got %d" % (len(fds), peer, len(received)))
        got[peer] = list(received)
    t.join()
    server.close()
    return got


def gather_directories(ranks, records):
    """All-gather of the per-rank block directories (numpy RECORD_DTYPE arrays): counts first, then the padded payload.
    Returns (records of all ranks sorted by (level, coord_id), owner rank of each)."""
    import numpy as np
    import torch
    if ranks.td is None:
        order = np.lexsort((records["coord_id"], records["level"]))
        return records[order], np.zeros(len(records), np.int32)
    dev = ranks.device if ranks.device is not None else torch.device("cpu")
    counts = torch.zeros(ranks.world, dtype=torch.int64, device=dev)
    mine = torch.tensor([len(records)], dtype=torch.int64, device=dev)
    ranks.td.all_gather_into_tensor(counts, mine)
    counts = counts.cpu().numpy()
    width = int(counts.max())
    item = records.dtype.itemsize
    payload = np.zeros(width * item, np.uint8)
    payload[:len(records) * item] = records.view(np.uint8).reshape(-1)
    out = torch.zeros(ranks.world * width * item, dtype=torch.uint8, device=dev)
    ranks.td.all_gather_into_tensor(out, torch.from_numpy(payload).to(dev))
    out = out.cpu().numpy().reshape(ranks.world, width * item)
    parts, owners = [], []
    for r in range(ranks.world):
        parts.append(out[r, :int(counts[r]) * item].copy().view(records.dtype))
        owners.append(np.full(int(counts[r]), r, np.int32))
    allrec, owner = np.concatenate(parts), np.concatenate(owners)
    order = np.lexsort((allrec["coord_id"], allrec["level"]))
    return allrec[order], owner[order]


class ShardedGrid:
    """One n^3 grid polygonized by all ranks (one GPU each).  Usage:
        sg = ShardedGrid(ranks, n)            # cube: local slab + peers' slabs mapped over NVLink
        d, m, b = sg.slab_tensors()           # torch uint8/int8 views [n/world, n, n] of the LOCAL slab: fill them
        sg.ready()                            # barrier: every slab is filled
        info = sg.polygonize()                # two phases around the one all-gather; this rank's blocks
        directory, owner = sg.directory()     # every rank's blocks (all-gather), reference order
    """

    def __init__(self, ranks, n, device_index=None, key=None, planes=None):
        import torch
        from . import capi
        self.ranks, self.n = ranks, n
        self.planes = None if planes is None else [int(p) for p in planes]  # None = equal slabs
        self.rank, self.world = ranks.rank, ranks.world
        self.device_index = ranks.local_rank if device_index is None else device_index
        self.device = torch.device("cuda", self.device_index)
        self.ctx = capi.Context(self.device_index)
        self.ctx.cube_create(n, self.rank, self.world, self.planes)
        if self.world > 1:
            fds = [self.ctx.cube_export(c) for c in range(3)]
            key = key or os.environ.get("MASTER_PORT", "0")
            got = exchange_fds(self.rank, self.world, fds, "%s-%d" % (key, n))
            for peer, theirs in got.items():
                for c, fd in enumerate(theirs):
                    self.ctx.cube_import(peer, c, fd)
                    os.close(fd)
            for fd in fds:
                os.close(fd)
        self._stream = torch.cuda.ExternalStream(self.ctx.exchange_stream(), device=self.device)

    def slab_tensors(self):
        import torch
        from . import capi
        d, m, b, size = self.ctx.cube_slab()
        z0, z1 = slab_planes(self.n, self.rank, self.world, self.planes)
        shape = (z1 - z0, self.n, self.n)
        view = lambda p: torch.as_tensor(capi.DevicePointer(p, size), device=self.device).view(shape)  # noqa: E731
        return view(d).view(torch.int8), view(m), view(b)

    def ready(self):
        import torch
        torch.cuda.synchronize(self.device)
        self.ranks.barrier()

    def polygonize(self, flags=0, max_attempts=4):
        import torch
        from . import capi
        td = self.ranks.td
        for _ in range(max_attempts):
            self.ctx.polygonize_sharded(self.rank, self.world, 0, flags, self.planes)
            x = self.ctx.shard_exchange_info(self.rank, self.world, self.planes)
            if self.world > 1 and x.pages_bytes:
                # the ONE data-path exchange, stream-ordered after phase 0 on the context's exchange stream, in place: the
                # material pages + flags of the last nested level.  Equal slabs: all-gather; unequal: rank 0 (the only
                # consumer, it builds the coarse levels) receives every other rank's range in one NCCL group.
                with torch.cuda.stream(self._stream):
                    pages = torch.as_tensor(capi.DevicePointer(x.pages, x.pages_bytes), device=self.device)
                    valid = torch.as_tensor(capi.DevicePointer(x.valid, x.valid_bytes), device=self.device)
                    if self.planes is None:
                        for full in (pages, valid):
                            chunk = full.numel() // self.world
                            td.all_gather_into_tensor(full, full[self.rank * chunk:(self.rank + 1) * chunk])
                    else:
                        def rng(r, unit):
                            return slice(self.planes[r] // x.layer_planes * x.layer_blocks * unit, self.planes[r + 1] // x.layer_planes * x.layer_blocks * unit)
                        ops = []
                        for full, unit in ((pages, 8192), (valid, 1)):
                            if self.rank == 0:
                                ops += [td.P2POp(td.irecv, full[rng(r, unit)], r) for r in range(1, self.world)]
                            else:
                                ops.append(td.P2POp(td.isend, full[rng(self.rank, unit)], 0))
                        for req in td.batch_isend_irecv(ops):
                            req.wait()
            rc = self.ctx.polygonize_sharded(self.rank, self.world, 1, flags, self.planes)
            # an overflow on any rank repeats the run on every rank (the exchange is collective)
            if self.ranks.max_over_ranks(1.0 if rc != 0 else 0.0) == 0.0:
                return self.ctx.info()
        raise capi.VxbError("sharded run: output arenas kept overflowing")

    def directory(self):
        import numpy as np
        from . import capi
        info = self.ctx.info()
        records = np.zeros(info.block_count, capi.RECORD_DTYPE)
        self.ctx._check(self.ctx.L.vxb_result_download(self.ctx.h, capi._ptr(records), None, None, None, None), "vxb_result_download")
        return gather_directories(self.ranks, records)

    def close(self):
        self.ctx.close()
