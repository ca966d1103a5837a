"""bench.py's reference arm (`--impl reference`) runs on CPU: its JSON line must keep the driver's contract."""
import json
import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line(reference):
    env = dict(os.environ, OMP_NUM_THREADS="1")  # what torchrun exports: the arm must still use every core it may
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--size", "64", "--steps", "2", "--warmup", "1"],
                         env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "Mvoxels/s polygonized" and line["unit"] == "Mvoxels/s"
    assert line["higher_is_better"] is True and line["steps"] == 2 and line["warmup"] == 1 and line["value"] > 0
    base = line["cpu_baseline"]
    assert base["kind"] == "reference" and base["value"] == line["value"] and base["cores"] == len(os.sched_getaffinity(0))
    assert line["e2e"] == {"value": line["value"], "unit": "Mvoxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "64^3" in line["config"]["workload"]


def test_product_arm_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--size", "64", "--steps", "1"], capture_output=True, text=True, timeout=600)
    assert out.returncode != 0                      # no CPU fallback: the product arm must not produce a number here
    assert not any(l.startswith("{") and '"value"' in l for l in out.stdout.splitlines())
