"""The CPU oracle against its own committed known-answer set, plus structural invariants.

There are no reference-owned golden vectors for this path: the digests under
tests/golden/ were produced by this oracle (make_golden.py) -- this test pins the oracle against
regressions and against the independently derived counts in SURVEY.md / BASELINE.md."""
import json
import os

import numpy as np
import pytest

import oracle
from tests.conftest import ROOT
from tests.support.digest import f32_equal, mesh_digests

GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "digests.json")))
SMALL = [k for k in sorted(GOLDEN) if int(k.rsplit("_", 3)[1][1:]) <= 128]


@pytest.mark.parametrize("key", SMALL)
def test_oracle_reproduces_digest(key):
    name, r_, b_, f_ = key.rsplit("_", 3)
    m = oracle.mesh_run(name, int(r_[1:]), float(b_[1:]), flags=int(f_[1:]))
    assert mesh_digests(m.positions, m.normals, m.keys, m.nibbles, m.quads, m.n_invalid_quads) == GOLDEN[key]
    m.free()


def test_survey_probe_counts():
    """vertex counts the survey derived independently with a numpy probe"""
    assert GOLDEN["torus_r128_b2_f0"]["n_vertices"] == 23248
    assert GOLDEN["mandelbulb_r128_b5_f0"]["n_vertices"] == 21208
    assert GOLDEN["mandelbulb_r256_b5_f0"]["n_vertices"] == 104232
    assert GOLDEN["p_key_r64_b2_f0"]["n_vertices"] == 64 * 64  # one plane z = -0.75


def test_npz_fixture_matches_oracle():
    g = np.load(os.path.join(ROOT, "tests", "golden", "torus_r32_b2.npz"))
    m = oracle.mesh_run("torus", 32, 2.0)
    assert np.array_equal(m.keys, g["keys"]) and np.array_equal(m.quads, g["quads"]) and np.array_equal(m.nibbles, g["nibbles"])
    assert f32_equal(m.positions, g["positions"]).all() and f32_equal(m.normals, g["normals"]).all()
    m.free()


def test_structure_invariants():
    m = oracle.mesh_run("mandelbulb", 64, 5.0)
    k = m.keys
    assert np.all(k[1:] > k[:-1]), "vertex order must be ascending key order (mesh.rs:224-226, SURVEY F10)"
    z = (k >> np.uint64(32)).astype(np.int64)
    assert z.min() >= 1 and z.max() <= 63, "faithful mode labels slices 1..R-1 (SURVEY F3)"
    assert m.quads.max() < len(k)
    # vertices lie inside their cell (within rounding)
    size = np.float32(5.0) / np.float32(63)
    x = (k & np.uint64(0xFFFF)).astype(np.float32)
    lo = np.float32(-2.5) + size * x
    assert np.all(m.positions[:, 0] >= lo - 1e-5) and np.all(m.positions[:, 0] <= lo + size + 1e-5)
    m.free()


def test_all_slices_mode_is_superset():
    a = oracle.mesh_run("torus", 32, 2.0)
    b = oracle.mesh_run("torus", 32, 2.0, flags=oracle.FLAG_ALL_SLICES)
    shift = np.uint64(1) << np.uint64(32)
    assert set((a.keys - shift).tolist()) <= set(b.keys.tolist())
    a.free()
    b.free()


def test_rust_f32_display():
    cases = {1.0: "1", 0.1: "0.1", -0.0: "-0", 1e30: "1000000000000000000000000000000", 1e-10: "0.0000000001",
             3.4028235e38: "340282350000000000000000000000000000000", 123456.789: "123456.79", 16777216.0: "16777216",
             0.30000001192092896: "0.3", -2.5: "-2.5", 1.17549435e-38: "0.000000000000000000000000000000000000011754944"}
    for v, s in cases.items():
        assert oracle.rust_f32(v) == s
    assert oracle.rust_f32(float("nan")) == "NaN" and oracle.rust_f32(float("inf")) == "inf" and oracle.rust_f32(float("-inf")) == "-inf"


def test_stl_ply_text_shape(tmp_path):
    m = oracle.mesh_run("torus", 16, 2.0)
    m.write_stl(tmp_path / "t.stl")
    m.write_ply(tmp_path / "t.ply")
    stl = open(tmp_path / "t.stl").read().split("\n")
    assert stl[0] == "solid" and stl[-2] == "endsolid" and stl[1].startswith("facet normal ") and stl[2] == "\touter loop"
    assert stl[3].startswith("\t\tvertex ") and stl[6] == "\tendloop" and stl[7] == "endfacet"
    assert (len(stl) - 3) == 7 * 2 * len(m.quads)
    ply = open(tmp_path / "t.ply").read().split("\n")
    assert ply[:3] == ["ply", "format ascii 1.0", "comment written by rust-sdf"]
    assert ply[3] == f"element vertex {len(m.keys)}" and ply[10] == f"element face {2 * len(m.quads)}" and ply[12] == "end_header"
    m.free()


def test_plugin_sdf_equals_hand_transcriptions(built):
    """oracle SDF id "plugin" (the front-end's emitted code compiled for the host) against the hand
    transcriptions in oracle/sdf_examples.h: the same meshes, bit for bit, for every example input"""
    from tests.conftest import load_example_shader
    from tests.support import host_eval
    from tests.support.digest import f32_equal
    for name, res, bounds in (("torus", 48, 2.0), ("martin_cube", 40, 2.0), ("p_key", 40, 20.0), ("mandelbulb", 40, 5.0)):
        oracle.set_plugin(host_eval.scalar_function(load_example_shader(name).lower_to_cuda()))
        a = oracle.mesh_run("plugin", res, bounds)
        b = oracle.mesh_run(name, res, bounds)
        assert len(a.keys) > 500
        assert np.array_equal(a.keys, b.keys) and np.array_equal(a.quads, b.quads) and np.array_equal(a.nibbles, b.nibbles)
        assert f32_equal(a.positions, b.positions).all() and f32_equal(a.normals, b.normals).all()
        a.free()
        b.free()
    oracle.set_plugin(None)


def test_slab_extract_of_a_whole_run_equals_the_slab_run():
    """the z-slab digest form used for grids too large for a full oracle run (tests/support/slabs.py, BASELINE config 5):
    extracting [z, z+2) from a whole-grid mesh gives the same digest as oracle.mesh_run(z_begin=z, z_end=z+2)"""
    from tests.support.slabs import extract_slab, slab_digest, stratified_pairs
    for name, res, bounds in (("mandelbulb", 96, 5.0), ("torus", 64, 2.0)):
        whole = oracle.mesh_run(name, res, bounds)
        qmax = whole.quads.max(axis=1)
        assert np.all(qmax[1:] >= qmax[:-1]), "a quad's largest index is its emitting vertex (mesh.rs:286-320)"
        n_checked = 0
        for z in stratified_pairs(res - 1, 12):
            part = oracle.mesh_run(name, res, bounds, z_begin=z, z_end=z + 2)
            want = slab_digest(part.keys, part.nibbles, part.positions, part.normals, part.quads)
            i0, i1 = extract_slab(whole.keys, 1, z, z + 2)
            j0, j1 = int(np.searchsorted(qmax, i0)), int(np.searchsorted(qmax, i1))
            got = slab_digest(whole.keys[i0:i1], whole.nibbles[i0:i1], whole.positions[i0:i1], whole.normals[i0:i1], whole.quads[j0:j1], index_base=i0)
            assert got == want, (name, z)
            n_checked += got["n_vertices"]
            part.free()
        assert n_checked > 0
        whole.free()


def test_slab_golden_file_is_well_formed():
    d = json.load(open(os.path.join(ROOT, "tests", "golden", "slab_digests.json")))
    assert "mandelbulb_r4096_b5_f0" in d and len(d["mandelbulb_r4096_b5_f0"]["slabs"]) >= 32
    assert sum(v["n_vertices"] for v in d["mandelbulb_r4096_b5_f0"]["slabs"].values()) > 500_000
