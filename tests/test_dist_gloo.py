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
print("RESULT rank=%d ms=%.1f tile_digest=%s verts=%d total=%d mvox=%.3f" % (r.rank, ms, digest[:12], verts, total_verts, whole_job_throughput(n, r.world, ms)))
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
