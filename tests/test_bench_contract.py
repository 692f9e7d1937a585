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


def test_parity_digest_of_slab_parts_equals_the_whole_mesh_digest(built):
    """bench.py's parity_check hashes the slabs of all ranks in z order, array by array; that must be the digest
    tests/support/digest.py gives for the assembled mesh (and what tests/golden/digests.json holds)"""
    import importlib.util
    import numpy as np
    import oracle
    from tests.support.digest import mesh_digests
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    o = oracle.mesh_run("torus", 64, 2.0)
    whole = mesh_digests(o.positions, o.normals, o.keys, o.nibbles, o.quads, o.n_invalid_quads)
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", "digests.json")))
    assert whole == golden[bench.golden_key("torus", 64, 2.0, 0)]
    # three "ranks": vertices cut at arbitrary places, quads cut elsewhere (each rank's quads already carry global indices);
    # one rank holds nothing at all
    nv, nq = len(o.keys), len(o.quads)
    cuts_v, cuts_q = [0, nv // 3, nv // 3, nv], [0, nq // 2, nq // 2, nq]
    pos = np.array(o.positions, np.float32)
    pos[5, 1] = np.float32("nan")     # NaN payloads are canonicalised the same way in both
    whole = mesh_digests(pos, o.normals, o.keys, o.nibbles, o.quads, o.n_invalid_quads)
    parts = []
    for k in range(3):
        v0, v1, q0, q1 = cuts_v[k], cuts_v[k + 1], cuts_q[k], cuts_q[k + 1]
        parts.append(({"keys": o.keys[v0:v1], "nibbles": o.nibbles[v0:v1], "quads": o.quads[q0:q1], "positions": pos[v0:v1], "normals": o.normals[v0:v1]},
                      o.n_invalid_quads if k == 0 else 0))
    assert bench.digest_parts(parts) == whole
    rec = bench.parity_record(bench.digest_parts(parts), "no_such_key")
    assert rec["golden"] is None and rec["ok"] is None
    o.free()
