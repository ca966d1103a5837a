"""The N>1 plumbing of bench.py (voxels_b200/dist.py) on CPU: world_size 2, gloo backend.
Each rank works on its own terrain tile; the only communication is a barrier and max/sum reductions."""
import os
import socket
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, time
sys.path.insert(0, os.environ["VXB_REPO"]); sys.path.insert(0, os.path.join(os.environ["VXB_REPO"], "tests"))
import numpy as np, torch
from voxels_b200 import synth
from voxels_b200.dist import Ranks, tile_origin, whole_job_throughput
import restate, golden_hash
r = Ranks("gloo", torch.device("cpu"))
assert r.world == 2
n = 32
dist, mat, blend = (t.numpy() for t in synth.terrain(n, "cpu", origin=tile_origin(r.rank, n)))
E = restate.Restate()
h = E.run(dist, mat, blend)
digest = golden_hash.level_digests(E.level(h, 0))["exact"]
verts = len(E.level(h, 0).verts)
r.barrier()
fake_ms = 10.0 + 5.0 * r.rank                      # rank 1 is "slower": the job time is the max over ranks
ms = r.max_over_ranks(fake_ms)
total_verts = r.sum_over_ranks(verts)
os.write(1, ("RESULT rank=%d ms=%.1f tile_digest=%s verts=%d total=%d mvox=%.3f\n" % (r.rank, ms, digest[:12], verts, total_verts, whole_job_throughput(n, r.world, ms))).encode())  # one atomic write per rank
r.close()
'''


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_ranks_gloo(tmp_path):
    if not os.path.exists(os.path.join(REPO, "build", "oracle", "libvxr_restate.so")):
        pytest.skip("restatement library not built")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, VXB_REPO=REPO, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(script)]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = sorted(l for l in out.stdout.splitlines() if l.startswith("RESULT"))
    assert len(lines) == 2, out.stdout + out.stderr
    fields = [dict(kv.split("=") for kv in l.split()[1:]) for l in lines]
    assert fields[0]["ms"] == fields[1]["ms"] == "15.0"                  # max over ranks, seen by both
    assert fields[0]["tile_digest"] != fields[1]["tile_digest"]          # independent tiles (different origins)
    assert fields[0]["total"] == fields[1]["total"] == str(int(fields[0]["verts"]) + int(fields[1]["verts"]))
    assert abs(float(fields[0]["mvox"]) - 2 * 32 ** 3 / 15e-3 / 1e6) < 1e-3  # whole-job throughput counts both tiles


SHARD_WORKER = r'''
import os, sys, tempfile
sys.path.insert(0, os.environ["VXB_REPO"])
import numpy as np, torch
from voxels_b200 import capi
from voxels_b200.dist import Ranks, exchange_fds, gather_directories, owned_pieces, default_group_planes
r = Ranks("gloo", torch.device("cpu"))
assert r.world == 2
# 1. descriptors travel between the ranks (what carries the cuMemExportToShareableHandle handles of the slabs)
mine = []
for c in range(3):
    f = tempfile.TemporaryFile()
    f.write(b"rank%d-channel%d" % (r.rank, c)); f.flush()
    mine.append(f)
got = exchange_fds(r.rank, r.world, [f.fileno() for f in mine], os.environ["MASTER_PORT"])
peer = 1 - r.rank
texts = []
for c, fd in enumerate(got[peer]):
    with os.fdopen(fd, "rb") as f:
        f.seek(0); texts.append(f.read().decode())
assert texts == ["rank%d-channel%d" % (peer, c) for c in range(3)], texts
# 2. directory all-gather: counts, then padded payload; reference order = (level, coord_id)
n = 1024
g = default_group_planes(n, r.world)
mine_pieces = owned_pieces(n, r.rank, r.world, g)
recs = np.zeros(3 + r.rank, capi.RECORD_DTYPE)
recs["level"] = [0, 0, 1] + [2] * r.rank
recs["coord_id"] = [10 + 100 * r.rank, 5 + 100 * r.rank, 7 + r.rank] + [0] * r.rank
recs["id"] = recs["coord_id"] + 1000 * recs["level"]
allrec, owner = gather_directories(r, recs)
assert len(allrec) == 7 and list(allrec["level"]) == sorted(allrec["level"])
keys = list(zip(allrec["level"].tolist(), allrec["coord_id"].tolist()))
assert keys == sorted(keys)
assert owner.tolist() == [0, 0, 1, 1, 0, 1, 1], owner.tolist()
# 3. the NCCL id of the in-step all-gathers travels as a python object
ids = [b"x" * 128 if r.rank == 0 else None]
r.td.broadcast_object_list(ids, src=0)
assert ids[0] == b"x" * 128
os.write(1, ("RESULT rank=%d group=%d pieces=%d first=%d-%d blocks=%d\n" % (r.rank, g, len(mine_pieces), mine_pieces[0][1], mine_pieces[0][2], len(allrec))).encode())
r.close()
'''


def test_sharded_plumbing_two_ranks_gloo(tmp_path):
    script = tmp_path / "shard_worker.py"
    script.write_text(SHARD_WORKER)
    env = dict(os.environ, VXB_REPO=REPO, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(script)]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = sorted(l for l in out.stdout.splitlines() if l.startswith("RESULT"))
    assert lines == ["RESULT rank=0 group=32 pieces=16 first=0-32 blocks=7", "RESULT rank=1 group=32 pieces=16 first=32-64 blocks=7"], out.stdout + out.stderr


def test_cyclic_pieces_cover_the_grid_once():
    sys.path.insert(0, REPO)
    from voxels_b200.dist import default_group_planes, owned_pieces
    for n, world in ((1024, 1), (1024, 2), (1024, 4), (1024, 8), (2048, 8), (512, 4), (256, 2)):
        g = default_group_planes(n, world)
        assert g % 32 == 0 and n % g == 0 and (n // g) % world == 0
        assert (g * n * n) % (2 << 20) == 0                        # allocation granularity of the volume pieces
        seen = []
        for r in range(world):
            mine = owned_pieces(n, r, world, g)
            assert all(p % world == r and z1 - z0 == g and z0 == p * g for p, z0, z1 in mine)
            seen += [p for p, _, _ in mine]
        assert sorted(seen) == list(range(n // g))                 # every piece exactly one owner
    # a terrain's surface (a few z-layers) is spread over all ranks' memories: 8 ranks, 1024^3 => 32-plane pieces
    assert default_group_planes(1024, 8) == 32 and default_group_planes(2048, 8) == 32


def test_parse_cpulist():
    sys.path.insert(0, REPO)
    from voxels_b200.dist import parse_cpulist
    assert parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert parse_cpulist("5") == [5] and parse_cpulist("") == []
