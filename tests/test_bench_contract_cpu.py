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
    assert base["kind"] == "reference" and base["value"] == line["value"]
    # every core it may use - or, in a container whose CPU quota is smaller, the quota if that ran faster (both were tried)
    assert str(base["cores"]) in base["thread_trials_s"] and str(len(os.sched_getaffinity(0))) in base["thread_trials_s"]
    assert base["cores"] == int(min(base["thread_trials_s"], key=base["thread_trials_s"].get))
    assert line["e2e"] == {"value": line["value"], "unit": "Mvoxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "64^3" in line["config"]["workload"]


def test_product_arm_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--size", "64", "--steps", "1"], capture_output=True, text=True, timeout=600)
    assert out.returncode != 0                      # no CPU fallback: the product arm must not produce a number here
    assert not any(l.startswith("{") and '"value"' in l for l in out.stdout.splitlines())


def test_cgroup_quota_and_throttling_are_reported(tmp_path):
    """bench.cgroup_cpu / cgroup_delta: what the CPU numbers carry about the container's CPU quota (cgroup v2 files)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(REPO, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec.loader.exec_module(bench)
    finally:
        sys.argv = argv
    assert bench.cgroup_cpu(str(tmp_path / "missing")) == {} and bench.cgroup_delta({}, {}) is None
    (tmp_path / "cpu.max").write_text("1600000 100000\n")
    (tmp_path / "cpu.stat").write_text("usage_usec 5\nnr_periods 100\nnr_throttled 7\nthrottled_usec 2500\n")
    before = bench.cgroup_cpu(str(tmp_path))
    assert before == {"quota_cpus": 16.0, "nr_periods": 100, "nr_throttled": 7, "throttled_usec": 2500}
    (tmp_path / "cpu.stat").write_text("usage_usec 9\nnr_periods 160\nnr_throttled 37\nthrottled_usec 1002500\n")
    assert bench.cgroup_delta(before, bench.cgroup_cpu(str(tmp_path))) == {"quota_cpus": 16.0, "throttled_periods": 30, "periods": 60, "throttled_ms": 1000.0}
    (tmp_path / "cpu.max").write_text("max 100000\n")
    assert bench.cgroup_cpu(str(tmp_path))["quota_cpus"] is None
