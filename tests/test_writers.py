"""Mesh file writers of the library (csrc/writers.cpp) on the CPU, through s2m_write_mesh_arrays:

* byte for byte the oracle's restatement of /root/reference/src/mesh.rs:8-48 (STLWriter), :50-141
  (PLYWriter), :167-210 (TriangleMesh::write_*_to_file) on oracle meshes,
* the float text against an independent formatter (numpy's shortest round-trip digits laid out the
  way Rust's `{}` lays them out: never an exponent, no trailing ".0"),
* z-slab parts (global vertex bases + halo positions) give the same file as the whole mesh,
* binary STL layout, extension handling and error returns.
"""
import struct

import numpy as np
import pytest

import oracle
import sdf2mesh_b200 as s2m


def rust_display_f32(x) -> str:
    """Rust `format!("{}", x)` for an f32, from numpy's shortest-unique digits (independent of both
    csrc/writers.cpp and oracle.cpp, which use std::to_chars)"""
    x = np.float32(x)
    if np.isnan(x):
        return "NaN"
    if np.isinf(x):
        return "inf" if x > 0 else "-inf"
    s = np.format_float_positional(x, unique=True, trim="-")
    if s in ("0", "-0"):
        return s
    return s[:-1] if s.endswith(".") else s


def triangles_of(quads):
    """lib.rs:199-204 Quad::make_triangles: (q2,q1,q0), (q0,q3,q2)"""
    q = np.asarray(quads, np.int64)
    return np.stack([q[:, [2, 1, 0]], q[:, [0, 3, 2]]], axis=1).reshape(-1, 3)


def expected_stl(positions, quads) -> str:
    """mesh.rs:8-48 STLWriter with Triangle::normal lib.rs:181-185, formatted by rust_display_f32"""
    f = rust_display_f32
    out = ["solid\n"]
    for t in triangles_of(quads):
        v0, v1, v2 = (positions[i].astype(np.float32) for i in t)
        a, b = (v2 - v0).astype(np.float32), (v1 - v0).astype(np.float32)
        with np.errstate(all="ignore"):
            n = np.array([np.float32(a[1] * b[2]) - np.float32(a[2] * b[1]), np.float32(a[2] * b[0]) - np.float32(a[0] * b[2]),
                          np.float32(a[0] * b[1]) - np.float32(a[1] * b[0])], np.float32)
        out.append(f"facet normal {f(n[0])} {f(n[1])} {f(n[2])}\n\touter loop\n")
        for v in (v0, v1, v2):
            out.append(f"\t\tvertex {f(v[0])} {f(v[1])} {f(v[2])}\n")
        out.append("\tendloop\nendfacet\n")
    out.append("endsolid\n")
    return "".join(out)


def expected_ply(positions, normals, quads) -> str:
    """mesh.rs:50-141 PLYWriter as TriangleMesh::write_ply_to_file drives it (mesh.rs:198-210)"""
    f = rust_display_f32
    tris = triangles_of(quads)
    out = ["ply\nformat ascii 1.0\ncomment written by rust-sdf\n", f"element vertex {len(positions)}\n",
           "property float x\nproperty float y\nproperty float z\nproperty float nx\nproperty float ny\nproperty float nz\n",
           f"element face {len(tris)}\nproperty list uchar int vertex_index\nend_header\n"]
    for p, n in zip(positions, normals):
        out.append(f"{f(p[0])} {f(p[1])} {f(p[2])} {f(n[0])} {f(n[1])} {f(n[2])}\n")
    for t in tris:
        out.append(f"3 {t[0]} {t[1]} {t[2]}\n")
    return "".join(out)


@pytest.mark.parametrize("name,res,bounds", [("torus", 24, 2.0), ("mandelbulb", 48, 5.0), ("p_key", 16, 20.0)])
def test_ascii_writers_equal_the_oracle_and_the_independent_text(built, tmp_path, name, res, bounds):
    o = oracle.mesh_run(name, res, bounds)
    assert len(o.quads) > 100
    for ext, quads in (("stl", o.quads), ("PLY", o.quads), ("Stl", o.quads.astype(np.uint32))):
        mine, ref = tmp_path / f"mine.{ext}", tmp_path / f"ref.{ext}"
        s2m.write_mesh_arrays([(o.positions, o.normals, quads)], mine)
        (o.write_ply if ext == "PLY" else o.write_stl)(ref)
        text = mine.read_bytes()
        assert text == ref.read_bytes()
        want = expected_ply(o.positions, o.normals, o.quads) if ext == "PLY" else expected_stl(o.positions, o.quads)
        assert text.decode() == want
    o.free()


def test_float_text_of_awkward_values(built, tmp_path):
    """every class of f32 the formatter distinguishes, as vertex coordinates of one quad"""
    rng = np.random.default_rng(7)
    bits = rng.integers(0, 2**32, 3 * 4 * 600, dtype=np.uint64).astype(np.uint32)
    vals = bits.view(np.float32).copy()
    special = np.array([0.0, -0.0, 1.0, -1.0, 0.1, 1e-10, 1e30, 3.4028235e38, -3.4028235e38, 1.17549435e-38, 1e-45, -1e-45, 16777216.0,
                        16777218.0, 123456.789, 0.30000001192092896, 9.999999e-5, 1e-4, 1e7, 1e8, 0.5, 1.5e-7, np.inf, -np.inf, np.nan], np.float32)
    vals[:special.size] = special
    pos = vals.reshape(-1, 3)
    quads = np.arange(pos.shape[0], dtype=np.uint64).reshape(-1, 4)
    nrm = pos[::-1].copy()
    with np.errstate(all="ignore"):
        want_stl, want_ply = expected_stl(pos, quads), expected_ply(pos, nrm, quads)
    s2m.write_mesh_arrays([(pos, nrm, quads)], tmp_path / "a.stl")
    s2m.write_mesh_arrays([(pos, nrm, quads)], tmp_path / "a.ply")
    got = (tmp_path / "a.stl").read_text().split("\n")
    for i, (g, w) in enumerate(zip(got, want_stl.split("\n"))):
        if "NaN" in w and "facet normal" in w:
            continue   # the sign of a NaN produced by arithmetic is not defined; Rust prints "NaN" either way, so do we
        assert g == w, i
    assert (tmp_path / "a.ply").read_text() == want_ply


def test_z_slab_parts_write_the_whole_mesh(built, tmp_path):
    """two z-slabs as a multi-GPU run leaves them (own vertices + global base + the halo slice below) ==
    the whole mesh; a slab alone is a valid STL of its own quads"""
    o = oracle.mesh_run("torus", 32, 2.0)
    z = (o.keys >> np.uint64(32)).astype(np.int64)
    cut = int(np.median(z))
    lower = z < cut
    n0 = int(lower.sum())
    emit = np.maximum.reduce([z[o.quads[:, k].astype(np.int64)] for k in range(4)])  # a quad belongs to the slab of the cell that emits it: the largest z of its four
    q_low, q_up = o.quads[emit < cut], o.quads[emit >= cut]
    assert len(q_low) and len(q_up) and len(q_low) + len(q_up) == len(o.quads)
    first_halo = int(q_up.min())
    assert first_halo < n0   # the upper slab refers to vertices of the slice below it
    halo = o.positions[first_halo:n0]
    whole, parts = tmp_path / "whole.stl", tmp_path / "parts.stl"
    s2m.write_mesh_arrays([(o.positions, o.normals, o.quads)], whole)
    s2m.write_mesh_arrays([(o.positions[:n0], o.normals[:n0], q_low), (o.positions[n0:], o.normals[n0:], q_up, n0)], parts)
    assert parts.read_bytes() == whole.read_bytes()
    s2m.write_mesh_arrays([(o.positions[:n0], o.normals[:n0], q_low), (o.positions[n0:], o.normals[n0:], q_up, n0)], tmp_path / "parts.ply")
    s2m.write_mesh_arrays([(o.positions, o.normals, o.quads)], tmp_path / "whole.ply")
    assert (tmp_path / "parts.ply").read_bytes() == (tmp_path / "whole.ply").read_bytes()
    # the upper slab on its own: needs its halo
    with pytest.raises(s2m.S2mError) as e:
        s2m.write_mesh_arrays([(o.positions[n0:], None, q_up, n0)], tmp_path / "up.stl")
    assert "outside the given parts" in str(e.value)
    s2m.write_mesh_arrays([(o.positions[n0:], None, q_up, n0, halo)], tmp_path / "up.stl")
    assert (tmp_path / "up.stl").read_text() == expected_stl(o.positions, q_up)
    with pytest.raises(s2m.S2mError) as e:   # PLY names vertices by index: whole mesh only
        s2m.write_mesh_arrays([(o.positions[n0:], o.normals[n0:], q_up, n0, halo)], tmp_path / "up.ply")
    assert "whole mesh" in str(e.value)
    o.free()


@pytest.mark.parametrize("threads,chunk", [(0, 9), (1, 5), (3, 7), (8, 1), (2, 1000)])
def test_chunk_pipeline_keeps_the_order(built, tmp_path, monkeypatch, threads, chunk):
    """the writers format chunks on worker threads into a ring of buffers and write them in order: any
    thread count and chunk size gives the file of the (serial) oracle writer"""
    o = oracle.mesh_run("torus", 24, 2.0)
    monkeypatch.setenv("S2M_WRITER_THREADS", str(threads))
    monkeypatch.setenv("S2M_WRITER_CHUNK", str(chunk))
    for ext in ("stl", "ply"):
        s2m.write_mesh_arrays([(o.positions, o.normals, o.quads)], tmp_path / f"a.{ext}")
        (o.write_ply if ext == "ply" else o.write_stl)(tmp_path / f"ref.{ext}")
        assert (tmp_path / f"a.{ext}").read_bytes() == (tmp_path / f"ref.{ext}").read_bytes()
    s2m.write_mesh_arrays([(o.positions, None, o.quads)], tmp_path / "b.stl", binary_stl=True)
    monkeypatch.delenv("S2M_WRITER_THREADS")
    monkeypatch.delenv("S2M_WRITER_CHUNK")
    s2m.write_mesh_arrays([(o.positions, None, o.quads)], tmp_path / "b0.stl", binary_stl=True)
    assert (tmp_path / "b.stl").read_bytes() == (tmp_path / "b0.stl").read_bytes()
    o.free()


def test_writer_reports_a_full_disk(built, tmp_path):
    import os
    if not os.path.exists("/dev/full"):
        pytest.skip("no /dev/full")
    link = tmp_path / "full.stl"
    os.symlink("/dev/full", link)
    o = oracle.mesh_run("torus", 16, 2.0)
    with pytest.raises(s2m.S2mError) as e:
        s2m.write_mesh_arrays([(o.positions, o.normals, o.quads)], link)
    assert e.value.kind == "IO"
    o.free()


def test_binary_stl_layout(built, tmp_path):
    o = oracle.mesh_run("torus", 16, 2.0)
    s2m.write_mesh_arrays([(o.positions, None, o.quads)], tmp_path / "b.stl", binary_stl=True)
    raw = (tmp_path / "b.stl").read_bytes()
    tris = triangles_of(o.quads)
    assert len(raw) == 84 + 50 * len(tris) and struct.unpack("<I", raw[80:84])[0] == len(tris)
    rec = np.frombuffer(raw[84:], np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("attr", "<u2")]))
    assert np.array_equal(rec["v"].view(np.uint32), o.positions[tris].view(np.uint32)) and not rec["attr"].any()
    a, b = o.positions[tris[:, 2]] - o.positions[tris[:, 0]], o.positions[tris[:, 1]] - o.positions[tris[:, 0]]
    n = np.cross(a.astype(np.float64), b.astype(np.float64))
    ln = np.linalg.norm(n, axis=1)
    ok = ln > 1e-12
    assert np.allclose(rec["n"][ok], (n[ok] / ln[ok, None]), atol=2e-3) and np.allclose(np.linalg.norm(rec["n"][ok], axis=1), 1.0, atol=1e-5)
    o.free()


def test_extension_and_errors(built, tmp_path, capfd):
    pos = np.zeros((4, 3), np.float32)
    quads = np.array([[0, 1, 2, 3]], np.uint64)
    s2m.write_mesh_arrays([(pos, pos, quads)], tmp_path / "mesh.obj")   # mesh.rs:193: logged, Ok(()), no file
    assert not (tmp_path / "mesh.obj").exists() and "Unknown file extension: OBJ" in capfd.readouterr().err
    s2m.write_mesh_arrays([(pos, pos, quads)], tmp_path / "noext")
    assert not (tmp_path / "noext").exists()
    s2m.write_mesh_arrays([(pos, pos, quads)], tmp_path / "UPPER.STL")
    assert (tmp_path / "UPPER.STL").read_text().startswith("solid\nfacet normal 0 0 0\n")
    with pytest.raises(s2m.S2mError) as e:
        s2m.write_mesh_arrays([(pos, pos, quads)], tmp_path / "no_such_dir" / "a.stl")
    assert e.value.kind == "IO"
    with pytest.raises(s2m.S2mError):
        s2m.write_mesh_arrays([(pos, pos, np.array([[0, 1, 2, 4]], np.uint64))], tmp_path / "bad.stl")   # index past the end
    with pytest.raises(s2m.S2mError):
        s2m.write_mesh_arrays([(pos, None, quads)], tmp_path / "n.ply")   # PLY writes normals
    # an empty mesh is a valid file
    s2m.write_mesh_arrays([(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32), np.zeros((0, 4), np.uint64))], tmp_path / "e.stl")
    assert (tmp_path / "e.stl").read_text() == "solid\nendsolid\n"
    s2m.write_mesh_arrays([(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32), np.zeros((0, 4), np.uint64))], tmp_path / "e.ply")
    assert (tmp_path / "e.ply").read_text() == expected_ply(np.zeros((0, 3)), np.zeros((0, 3)), np.zeros((0, 4), np.uint64))
