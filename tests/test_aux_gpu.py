"""GPU tests for the pieces around the four kernels: device math == host math bit for bit, writers
byte-identical to the oracle's restatement of mesh.rs, candidate classification, cost probe."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import oracle
import sdf2mesh_b200 as s2m
from tests.conftest import ROOT, load_example_shader
from tests.support.digest import f32_equal

pytestmark = pytest.mark.gpu

# sdf3d(p): p.z selects the function, p.x / p.y are the operands -- runs s2m_math.h on the device
MATH_CUDA = open(os.path.join(ROOT, "tests", "support", "math_dispatch.h")).read().replace('#include "s2m_math.h"', "").replace("#pragma once", "") + """
S2M_HD float sdf3d(vec3 p) {
  const int fn = (int)p.z;
  return fn >= 100 ? s2m_dispatch2(fn, p.x, p.y) : s2m_dispatch1(fn, p.x);
}
"""


@pytest.fixture(scope="module")
def host_math(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("math") / "libmath_host.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-mfma", "-fPIC", "-shared", "-I", os.path.join(ROOT, "sdf2mesh_b200", "csrc"),
                           os.path.join(ROOT, "tests", "support", "math_host.cpp"), "-o", so])
    return ctypes.CDLL(so)


def test_device_math_is_bit_identical_to_host_math(ctx, host_math):
    """the premise of bit-exact parity: every pinned function gives the same bits on sm_100a and x86"""
    mod = s2m.Sdf3DShader.from_source(MATH_CUDA, s2m.SRC_CUDA).create_shader_module(ctx)
    rng = np.random.default_rng(11)
    n = 200_000
    special = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1.0, -1.0, 0.5, 2.0, 1e-45, -1e-45, 1.17549435e-38, 3.4028235e38, 88.72, 88.73,
                        -103.9, -104.1, 105615.0, 105616.0, 1e9, 1e30, 0.70710678, 1.41421356], np.float32)
    xs = np.concatenate([special, rng.uniform(-30, 30, n), 10 ** rng.uniform(-45, 38, n), -(10 ** rng.uniform(-45, 38, n)),
                         rng.uniform(-1, 1, n), rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32).view(np.float32)]).astype(np.float32)
    for fn in range(19):   # 19-21 (asinh, acosh, atanh; added after the round's last GPU run) join once they have run on a device
        pts = np.stack([xs, np.zeros_like(xs), np.full_like(xs, fn)], 1)
        dev = mod.eval_points(pts)
        host = np.empty_like(xs)
        host_math.s2m_host_map1(fn, xs.ctypes.data_as(ctypes.c_void_p), host.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(xs.size))
        eq = f32_equal(dev, host)
        assert eq.all(), f"fn {fn}: {np.count_nonzero(~eq)} mismatches, first at x={xs[~eq][0]!r}: dev={dev[~eq][0]!r} host={host[~eq][0]!r}"
    a = np.concatenate([np.repeat(special, len(special)), rng.uniform(-4, 4, n), 10 ** rng.uniform(-20, 20, n)]).astype(np.float32)
    b = np.concatenate([np.tile(special, len(special)), rng.uniform(-10, 10, n), rng.uniform(-3, 3, n)]).astype(np.float32)
    for fn in range(100, 108):
        pts = np.stack([a, b, np.full_like(a, fn)], 1)
        dev = mod.eval_points(pts)
        host = np.empty_like(a)
        host_math.s2m_host_map2(fn, a.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p), host.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(a.size))
        eq = f32_equal(dev, host)
        assert eq.all(), f"fn {fn}: {np.count_nonzero(~eq)} mismatches, first at ({a[~eq][0]!r}, {b[~eq][0]!r}): dev={dev[~eq][0]!r} host={host[~eq][0]!r}"


@pytest.mark.parametrize("name,res,bounds", [("torus", 32, 2.0), ("mandelbulb", 64, 5.0)])
def test_writers_byte_identical(ctx, tmp_path, name, res, bounds):
    """ASCII STL / PLY from the library == the oracle's restatement of mesh.rs, byte for byte"""
    p, _ = s2m.params_from_cli(res, bounds)
    r = s2m.mesh_run(ctx, load_example_shader(name).create_shader_module(ctx), p)
    o = oracle.mesh_run(name, res, bounds)
    for ext in ("stl", "PLY"):
        mine, ref = tmp_path / f"mine.{ext}", tmp_path / f"ref.{ext}"
        r.write_mesh(mine)
        (o.write_stl if ext == "stl" else o.write_ply)(ref)
        assert mine.read_bytes() == ref.read_bytes()
    r.write_mesh(tmp_path / "mesh.xyz")  # unknown extension: logged, not an error, no file (mesh.rs:193)
    assert not (tmp_path / "mesh.xyz").exists()
    r.write_stl_binary(tmp_path / "b.stl")
    raw = (tmp_path / "b.stl").read_bytes()
    nq = len(r.data().quads)
    assert len(raw) == 84 + 50 * 2 * nq and int.from_bytes(raw[80:84], "little") == 2 * nq
    tri0 = np.frombuffer(raw[84:84 + 48], np.float32).reshape(4, 3)
    q0 = r.data().quads[0]
    assert f32_equal(tri0[1], r.data().positions[q0[2]]).all() and f32_equal(tri0[3], r.data().positions[q0[0]]).all()
    r.free()
    o.free()


def test_candidates_are_a_superset_and_match_numpy(ctx):
    """K2+K3: candidate set == cells whose 8 slab corners are not all > tau / all < -tau, in linear order"""
    res, bounds = 40, 2.0
    bmin, bmax = oracle.cube_bounds(bounds)
    p = s2m.make_params(res, bmin, bmax, flags=s2m.MESH_KEEP_CANDIDATES)
    mod = load_example_shader("torus").create_shader_module(ctx)
    r = s2m.mesh_run(ctx, mod, p)
    d = r.data()
    slab = np.stack([s2m.debug_slab_plane(ctx, mod, p, z) for z in range(res + 1)])  # [z][y][x]
    size = (np.float32(bmax[0]) - np.float32(bmin[0])) / np.float32(res - 1)
    # default band (engine.cpp): max(1/16, 256 * sqrt(3) * ulp-of-coordinate-in-voxels) voxels
    ulp_voxels = np.float32(max(abs(bmin[0]), abs(bmax[0]))) * np.float32(1.1920929e-7) / size
    tau = np.float32(max(np.float32(0.0625), np.float32(256.0) * np.float32(1.7320508) * ulp_voxels)) * size
    P, N = slab > tau, slab < -tau

    def all8(m):
        return (m[:-1, :-1, :-1] & m[:-1, :-1, 1:] & m[:-1, 1:, :-1] & m[:-1, 1:, 1:] & m[1:, :-1, :-1] & m[1:, :-1, 1:] & m[1:, 1:, :-1] & m[1:, 1:, 1:])

    cand = ~(all8(P) | all8(N))
    cand = cand[:res - 1]  # faithful mode scans slices 0..res-2
    z, y, x = np.nonzero(cand)
    want = x.astype(np.uint64) | (y.astype(np.uint64) << np.uint64(16)) | (z.astype(np.uint64) << np.uint64(32))
    assert np.array_equal(d.candidates, want)
    vert = (d.keys - (np.uint64(1) << np.uint64(32)))
    assert np.isin(vert, d.candidates).all()
    keys, quads = d.keys.copy(), d.quads.copy()
    r.free()
    # an explicit band: wider band, more candidates, same mesh
    p.tau_voxels = 0.5
    r2 = s2m.mesh_run(ctx, mod, p)
    d2 = r2.data()
    assert d2.n_candidates > len(want) and np.array_equal(d2.keys, keys) and np.array_equal(d2.quads, quads)
    r2.free()


def test_cost_probe(ctx):
    p, _ = s2m.params_from_cli(256, 5.0)
    cost = s2m.cost_probe(ctx, load_example_shader("mandelbulb").create_shader_module(ctx), p, 32)
    assert len(cost) == 32 and np.all(cost > 0)
    assert cost[12:20].mean() > 2 * cost[[0, 1, 30, 31]].mean(), "the mandelbulb's cost is concentrated around z = 0"


def test_cli_binary_end_to_end(ctx, tmp_path):
    """the sdf2mesh executable with the reference's flags: same STL bytes as the oracle, same log lines"""
    exe = os.path.join(ROOT, "sdf2mesh_b200", "sdf2mesh")
    if not os.path.exists(exe):
        pytest.skip("CLI not built")
    out = tmp_path / "torus.stl"
    dbg = tmp_path / "debug.wgsl"
    p = subprocess.run([exe, "--sdf", os.path.join(ROOT, "examples", "torus.sdf3d"), "--resolution", "30", "--bounds", "2", "--mesh", str(out),
                        "--debug-wgsl", str(dbg), "--stats"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert "Reading SDF from" in p.stderr and "Resolution should be a power of 2 (actual resolution : 32)" in p.stderr
    assert "INFO  sdf2mesh] sdf3d::*" in p.stderr and "Mesh written to" in p.stderr
    o = oracle.mesh_run("torus", 32, 2.0)
    assert f"Mesh has {len(o.keys)} vertices." in p.stderr
    ref = tmp_path / "ref.stl"
    o.write_stl(ref)
    assert out.read_bytes() == ref.read_bytes()
    assert "fn sdf3d_torus(" in dbg.read_text() and "fn sdf3d(p: vec3f)" in dbg.read_text()
    # GLSL input, PLY output, short flags
    ply = tmp_path / "m.ply"
    p = subprocess.run([exe, "--glsl", os.path.join(ROOT, "examples", "mandelmesh.frag"), "-r", "64", "-b", "5", "-0", str(ply)], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    o2 = oracle.mesh_run("mandelbulb", 64, 5.0)
    refp = tmp_path / "ref.ply"
    o2.write_ply(refp)
    assert ply.read_bytes() == refp.read_bytes()
    # missing required flag / bad GLSL sdf name
    assert subprocess.run([exe, "-i", "x.sdf3d"], capture_output=True).returncode == 2
    p = subprocess.run([exe, "--glsl", os.path.join(ROOT, "examples", "mandelmesh.frag"), "--glsl-sdf", "nope", "-0", str(ply)], capture_output=True, text=True)
    assert p.returncode == 101 and "Missing SDF function" in p.stderr
    o.free()
    o2.free()


@pytest.mark.parametrize("name,res,bounds,splits", [("torus", 32, 2.0, [0, 9, 20, 31]), ("mandelbulb", 64, 5.0, [0, 30, 31, 50, 63])])
def test_z_slab_parts_write_the_same_file(ctx, tmp_path, name, res, bounds, splits):
    """s2m_write_mesh_parts over z-slab results (what a multi-GPU job holds) == the single-result file,
    and an STL written from one slab alone (its halo supplies the vertices below) is that slab's
    facets of the whole file"""
    p, _ = s2m.params_from_cli(res, bounds)
    m = load_example_shader(name).create_shader_module(ctx)
    full = s2m.mesh_run(ctx, m, p)
    parts, base = [], 0
    for zb, ze in zip(splits[:-1], splits[1:]):
        p.z_begin, p.z_end = zb, ze
        r = s2m.mesh_begin(ctx, m, p)
        n_own = r.info().n_vertices
        r.finish(base)
        assert r.info().global_vertex_base == base
        base += n_own
        parts.append(r)
    for ext in ("stl", "ply"):
        a, b = tmp_path / f"full.{ext}", tmp_path / f"parts.{ext}"
        full.write_mesh(a)
        s2m.write_mesh_parts(parts, b)
        assert a.read_bytes() == b.read_bytes()
    s2m.write_mesh_parts(parts, tmp_path / "pb.stl", binary_stl=True)
    full.write_stl_binary(tmp_path / "fb.stl")
    assert (tmp_path / "pb.stl").read_bytes() == (tmp_path / "fb.stl").read_bytes()
    # each slab on its own: facet text concatenates to the body of the whole file
    whole = (tmp_path / "full.stl").read_text().splitlines(keepends=True)
    body = []
    for k, r in enumerate(parts):
        f = tmp_path / f"slab{k}.stl"
        r.write_mesh(f)
        lines = f.read_text().splitlines(keepends=True)
        assert lines[0] == whole[0] and lines[-1] == whole[-1]
        body += lines[1:-1]
    assert body == whole[1:-1]
    with pytest.raises(s2m.S2mError):  # a PLY needs every vertex: one interior slab is not a mesh
        s2m.write_mesh_parts(parts[1:2], tmp_path / "bad.ply")
    for r in parts:
        r.free()
    full.free()


def test_cli_multi_gpu_threads(ctx, tmp_path):
    """sdf2mesh --gpus N (one host thread + context per GPU, in-process count exchange): same bytes"""
    import torch
    exe = os.path.join(ROOT, "sdf2mesh_b200", "sdf2mesh")
    n = torch.cuda.device_count()
    if not os.path.exists(exe) or n < 2:
        pytest.skip("needs the CLI and 2 GPUs")
    n = min(n, 4)
    one, many = tmp_path / "one.ply", tmp_path / "many.ply"
    args = [exe, "--glsl", os.path.join(ROOT, "examples", "mandelmesh.frag"), "-r", "128", "-b", "5"]
    assert subprocess.run(args + ["-0", str(one)], capture_output=True).returncode == 0
    p = subprocess.run(args + ["-0", str(many), "--gpus", str(n), "--stats"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert one.read_bytes() == many.read_bytes()


def test_module_compiled_without_a_device_then_instantiated(ctx):
    """s2m_module_instantiate: NVRTC output produced with ctx = NULL (what the CLI compiles while its
    context comes up, and what --gpus N compiles once for all GPUs) loaded into a context: the same
    mesh as the oracle, the same kernels as a module compiled for the context directly"""
    compiled = load_example_shader("torus").create_shader_module(None)
    with pytest.raises(s2m.S2mError):
        s2m.mesh_run(ctx, compiled, s2m.params_from_cli(32, 2.0)[0])   # not loaded anywhere
    m = compiled.instantiate(ctx)
    assert m.cubins() == compiled.cubins() and m.packed == compiled.packed and m.compile_ms()[2] > 0
    del compiled   # the instance owns copies of the cubins
    p, _ = s2m.params_from_cli(32, 2.0)
    r = s2m.mesh_run(ctx, m, p)
    o = oracle.mesh_run("torus", 32, 2.0)
    d = r.data()
    assert np.array_equal(d.keys, o.keys) and np.array_equal(d.quads, o.quads) and f32_equal(d.positions, o.positions).all()
    pts = np.random.default_rng(1).uniform(-1, 1, (1000, 3)).astype(np.float32)
    assert f32_equal(m.eval_points(pts), oracle.eval_points("torus", pts)).all()
    r.free()
    o.free()
