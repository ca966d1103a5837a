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
from voxels_b200.dist import Ranks, exchange_fds, gather_directories, slab_planes, split_level
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
n = 256
z0, z1 = slab_planes(n, r.rank, r.world)
recs = np.zeros(3 + r.rank, capi.RECORD_DTYPE)
recs["level"] = [0, 0, 1] + [2] * r.rank
recs["coord_id"] = [10 + 100 * r.rank, 5 + 100 * r.rank, 7 + r.rank] + [0] * r.rank
recs["id"] = recs["coord_id"] + 1000 * recs["level"]
allrec, owner = gather_directories(r, recs)
assert len(allrec) == 7 and list(allrec["level"]) == sorted(allrec["level"])
keys = list(zip(allrec["level"].tolist(), allrec["coord_id"].tolist()))
assert keys == sorted(keys)
assert owner.tolist() == [0, 0, 1, 1, 0, 1, 1], owner.tolist()
os.write(1, ("RESULT rank=%d slab=%d-%d split=%d blocks=%d\n" % (r.rank, z0, z1, split_level(n, r.world), len(allrec))).encode())
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
    assert lines == ["RESULT rank=0 slab=0-128 split=4 blocks=7", "RESULT rank=1 slab=128-256 split=4 blocks=7"], out.stdout + out.stderr


def test_split_level_matches_the_reference_block_sizes():
    sys.path.insert(0, REPO)
    from voxels_b200.dist import split_level
    assert split_level(2048, 8) == 5      # 16,32,64,128,256-voxel blocks nest in 256-plane slabs
    assert split_level(1024, 8) == 4
    assert split_level(1024, 1) == 7      # one rank: every level nests, nothing to exchange
    assert split_level(64, 2) == 2


def test_balanced_planes_even_out_the_work():
    sys.path.insert(0, REPO)
    import numpy as np
    from voxels_b200.dist import balanced_planes, layer_weights_from_directory, split_level
    from voxels_b200 import capi
    # a terrain: all the work in a few middle layers (z is up)
    w = np.zeros(128)
    w[50:70] = 100.0
    planes = balanced_planes(w, 4, 2)                       # boundaries on multiples of 2 layers = 32 planes
    assert planes[0] == 0 and planes[-1] == 2048 and all(p % 32 == 0 for p in planes) and planes == sorted(planes)
    loads = [w[a // 16:b // 16].sum() for a, b in zip(planes[:-1], planes[1:])]
    assert max(loads) <= 600.0                               # 2000 units over 4 ranks, cells of 200: the optimum is 600
    assert all(b > a for a, b in zip(planes[:-1], planes[1:]))   # no empty slab
    assert split_level(2048, 4, planes) == 2                 # 32-plane alignment: levels 0 and 1 nest
    assert split_level(2048, 4, [0, 512, 1024, 1536, 2048]) == 6
    # weights from a directory: a level-1 block spreads over the two level-0 layers it covers, plus the scan's volume term
    recs = np.zeros(2, capi.RECORD_DTYPE)
    recs["level"] = [0, 1]
    recs["coord_id"] = [3 * 16 * 16, 1 * 8 * 8]              # n = 256: level-0 block layer 3, level-1 block layer 1 (= level-0 layers 2, 3)
    recs["vertex_count"] = [100, 60]
    lw = layer_weights_from_directory(256, recs, scan_units_per_byte=0.0)
    assert lw.tolist() == [0.0, 0.0, 30.0, 130.0] + [0.0] * 12
