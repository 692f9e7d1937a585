"""bench.py's output contract on the CPU: the reference arm (the CPU port of the reference path, the
only leg that runs without a GPU) prints one JSON line with every key the driver reads; the product arm
refuses to run without a device instead of falling back to anything."""
import json
import os
import subprocess
import sys

import pytest

from tests.conftest import ROOT


def run_bench(*args, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600,
                          env={**os.environ, **(env or {})})


def test_reference_arm_line(built):
    p = run_bench("--impl", "reference", "--workload", "torus128", "--steps", "2", "--warmup", "1", "--ref-slices", "8")
    assert p.returncode == 0, p.stderr
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("Gvoxels/s") and d["unit"] == "Gvoxel/s"
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["dtype"] == "f32" and d["scaling"] in ("weak", "strong")
    assert d["config"]["workload"] == "torus128" and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "z-slices" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "Gvoxel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_only_rank0_works_under_torchrun(built):
    """N > 1: rank 0 alone runs and prints the line; the other ranks exit 0 without work"""
    p = run_bench("--impl", "reference", "--workload", "torus128", "--steps", "1", "--warmup", "0", "--ref-slices", "4", "--gpus", "2",
                  env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2", "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "29999"})
    assert p.returncode == 0, p.stderr
    assert not [l for l in p.stdout.splitlines() if l.startswith("{")]


def test_product_arm_needs_a_device(built):
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    p = run_bench("--workload", "torus128", "--steps", "1", "--warmup", "1", "--no-cpu-baseline")
    assert p.returncode != 0
    assert not [l for l in p.stdout.splitlines() if l.startswith("{")]   # no number without the CUDA path
