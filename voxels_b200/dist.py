"""torch.distributed plumbing for the multi-GPU runs (one process per GPU; NCCL on the B200 box, gloo in the CPU tests).

The polygonizer shards spatially with NO data-path collective: every rank owns an independent n^3 tile of the terrain
(`tile_origin`), so the only communication is the barrier around the timed region and the max-reduction of the per-rank
device times (the contract of bench.py)."""
import os


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
