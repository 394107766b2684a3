"""bench.py contract checks that run without a GPU: the reference arm (the oracle port timed on the host cores) prints
one JSON line with the keys the driver reads, and the product arm refuses to run without a device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, timeout=300):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT, capture_output=True,
                          text=True, timeout=timeout)


def test_reference_arm_json_line():
    r = _run(["--impl", "reference", "--steps", "2", "--warmup", "1", "--ref-frames", "20000"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference"
    assert j["unit"] == "frames/s" and j["higher_is_better"] is True and j["value"] > 0
    assert j["n_gpus"] == 1 and j["steps"] == 2 and j["warmup"] == 1
    assert j["dtype"] == "f32" and j["data"] == "synthetic" and j["vs_baseline"] is None
    assert "cfg2" in j["config"]["workload"] and j["config"]["k"] == 1000 and j["config"]["d"] == 10
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == j["value"]
    assert "parity unpinned" in cb["sample"]
    e = j["e2e"]
    assert e["value"] == j["value"] and e["unit"] == j["unit"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert abs(j["ms_per_step"] * 1e-3 * j["value"] - 20000) <= 1e-6 * 20000


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a machine without a GPU")
def test_product_arm_needs_a_gpu():
    r = _run(["--steps", "1", "--warmup", "3", "--no-cpu-baseline"])
    assert r.returncode != 0
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_under_torchrun_prints_one_line():
    """the driver launches the reference arm like the product arm (torchrun, one process per GPU): rank 0 alone runs it
    and prints the line, the other ranks exit 0 without work"""
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "bench.py"),
                        "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--ref-frames", "20000"],
                       cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["n_gpus"] == 2 and j["value"] > 0
