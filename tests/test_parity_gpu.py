"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the golden digests.

Bar: bit-exact keys (active cell set + order), sign nibbles, quad connectivity and invalid-quad
count; vertex positions and normals bit-exact as well (NaN == NaN), which is stricter than the
1e-4 voxel tolerance north_star asks for.
"""
import json
import os

import numpy as np
import pytest

import oracle
import sdf2mesh_b200 as s2m
from tests.conftest import ROOT, load_example_shader
from tests.support.digest import f32_equal, mesh_digests

pytestmark = pytest.mark.gpu

GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "digests.json")))
_modules = {}


def module_for(ctx, name):
    if name not in _modules:
        _modules[name] = load_example_shader(name).create_shader_module(ctx)
    return _modules[name]


def assert_same(d, o, what=""):
    assert len(d.keys) == len(o.keys), f"{what}: vertex count {len(d.keys)} vs oracle {len(o.keys)}"
    assert np.array_equal(d.keys, o.keys), f"{what}: active cell set / order differs"
    assert np.array_equal(d.nibbles, o.nibbles), f"{what}: sign nibbles differ"
    assert d.n_invalid_quads == o.n_invalid_quads, f"{what}: invalid quads {d.n_invalid_quads} vs {o.n_invalid_quads}"
    assert np.array_equal(d.quads, o.quads), f"{what}: quad connectivity differs"
    eq = f32_equal(d.positions, o.positions)
    assert eq.all(), f"{what}: {np.count_nonzero(~eq)} position components differ"
    eq = f32_equal(d.normals, o.normals)
    assert eq.all(), f"{what}: {np.count_nonzero(~eq)} normal components differ"


CASES = [("torus", 32, 2.0), ("torus", 128, 2.0), ("martin_cube", 64, 2.0), ("martin_cube", 128, 2.0),
         ("p_key", 64, 2.0), ("p_key", 128, 20.0), ("mandelbulb", 64, 5.0), ("mandelbulb", 128, 5.0)]


@pytest.mark.parametrize("name,res,bounds", CASES)
def test_mesh_matches_oracle(ctx, name, res, bounds):
    p, _ = s2m.params_from_cli(res, bounds)
    r = s2m.mesh_run(ctx, module_for(ctx, name), p)
    o = oracle.mesh_run(name, res, bounds)
    try:
        assert_same(r.data(), o, f"{name} {res}^3")
    finally:
        r.free()
        o.free()


@pytest.mark.parametrize("key", sorted(GOLDEN))
def test_mesh_matches_golden_digest(ctx, key):
    name, r_, b_, f_ = key.rsplit("_", 3)
    if name == "naga_sphere":
        pytest.skip("covered by tests/test_frontend_gpu.py")
    res, bounds, flags = int(r_[1:]), float(b_[1:]), int(f_[1:])
    p, _ = s2m.params_from_cli(res, bounds, flags=s2m.MESH_ALL_SLICES if flags & 1 else 0)
    r = s2m.mesh_run(ctx, module_for(ctx, name), p)
    d = r.data()
    got = mesh_digests(d.positions, d.normals, d.keys, d.nibbles, d.quads, d.n_invalid_quads)
    r.free()
    assert got == GOLDEN[key]


def test_golden_npz_fixture(ctx):
    g = np.load(os.path.join(ROOT, "tests", "golden", "torus_r32_b2.npz"))
    p, _ = s2m.params_from_cli(32, 2.0)
    r = s2m.mesh_run(ctx, module_for(ctx, "torus"), p)
    d = r.data()
    assert np.array_equal(d.keys, g["keys"]) and np.array_equal(d.quads, g["quads"]) and np.array_equal(d.nibbles, g["nibbles"])
    assert f32_equal(d.positions, g["positions"]).all() and f32_equal(d.normals, g["normals"]).all()
    r.free()


@pytest.mark.parametrize("name,bounds", [("torus", 2.0), ("martin_cube", 2.5), ("p_key", 20.0), ("mandelbulb", 5.0)])
def test_sdf_values_bit_exact(ctx, name, bounds):
    """device SDF == hand-transcribed oracle SDF, bit for bit, on random points"""
    rng = np.random.default_rng(7)
    pts = rng.uniform(-bounds / 2, bounds / 2, (300000, 3)).astype(np.float32)
    dev = module_for(ctx, name).eval_points(pts)
    ref = oracle.eval_points(name, pts)
    eq = f32_equal(dev, ref)
    assert eq.all(), f"{np.count_nonzero(~eq)} of {len(pts)} values differ"


def test_slab_plane_matches_oracle(ctx):
    """K1: corner values at the variant-A coordinates fl(bmin + fl(size*j))"""
    res, bounds = 100, 2.0  # not a power of two: exercises the padded pitch
    bmin, bmax = oracle.cube_bounds(bounds)
    p = s2m.make_params(res, bmin, bmax)
    size = (np.float32(bmax[0]) - np.float32(bmin[0])) / np.float32(res - 1)
    coords = (np.float32(bmin[0]) + size * np.arange(res + 1, dtype=np.float32)).astype(np.float32)
    for plane in (0, 37, res):
        got = s2m.debug_slab_plane(ctx, module_for(ctx, "torus"), p, plane)
        X, Y = np.meshgrid(coords, coords)
        pts = np.stack([X, Y, np.full_like(X, coords[plane])], -1).reshape(-1, 3)
        ref = oracle.eval_points("torus", pts).reshape(res + 1, res + 1)
        assert f32_equal(got, ref).all()


def test_exact_dense_mode_equals_default(ctx):
    """reference-cost mode (every cell evaluated with the reference's 8 corners) == candidate mode"""
    for name, res, bounds in [("torus", 48, 2.0), ("mandelbulb", 64, 5.0)]:
        p, _ = s2m.params_from_cli(res, bounds)
        a = s2m.mesh_run(ctx, module_for(ctx, name), p)
        p.flags = s2m.MESH_EXACT_DENSE
        b = s2m.mesh_run(ctx, module_for(ctx, name), p)
        da, db = a.data(), b.data()
        n = p.dims[0]  # the CLI rounds the resolution up to a power of two (main.rs:142-148)
        assert db.n_candidates == n * n * (n - 1)
        assert np.array_equal(da.keys, db.keys) and np.array_equal(da.quads, db.quads)
        assert f32_equal(da.positions, db.positions).all()
        a.free()
        b.free()


def test_all_slices_flag(ctx):
    p, _ = s2m.params_from_cli(64, 2.0, flags=s2m.MESH_ALL_SLICES)
    r = s2m.mesh_run(ctx, module_for(ctx, "torus"), p)
    o = oracle.mesh_run("torus", 64, 2.0, flags=oracle.FLAG_ALL_SLICES)
    assert_same(r.data(), o, "torus all-slices")
    r.free()
    o.free()


def test_non_cubic_grid_and_small_chunks(ctx):
    """dims differ per axis, resolution not a multiple of 32, slab forced into many z-chunks"""
    dims = (70, 45, 33)
    bmin = np.array([-1.0, -0.6, -0.5], np.float32)
    bmax = np.array([1.0, 0.7, 0.5], np.float32)
    o = oracle.mesh_run("torus", dims, bmin=bmin, bmax=bmax)
    for budget in (0, 71 * 46 * 4 * 3 + 4096):  # default, and ~3 planes per chunk
        p = s2m.make_params(dims, bmin, bmax, slab_budget_bytes=budget)
        r = s2m.mesh_run(ctx, module_for(ctx, "torus"), p)
        d = r.data()
        if budget:
            assert d.timings["chunks"] > 4
        assert_same(d, o, f"non-cubic budget={budget}")
        r.free()
    o.free()


def test_empty_and_degenerate(ctx):
    """no surface inside the box -> empty mesh; tiny grid works"""
    p = s2m.make_params(16, [5, 5, 5], [6, 6, 6])
    r = s2m.mesh_run(ctx, module_for(ctx, "torus"), p)
    d = r.data()
    assert len(d.keys) == 0 and len(d.quads) == 0 and d.n_invalid_quads == 0
    r.free()
    p, _ = s2m.params_from_cli(2, 2.0)
    r = s2m.mesh_run(ctx, module_for(ctx, "torus"), p)
    o = oracle.mesh_run("torus", 2, 2.0)
    assert_same(r.data(), o, "2^3")
    r.free()
    o.free()


@pytest.mark.parametrize("name,res,bounds,splits", [("torus", 64, 2.0, [0, 20, 31, 50, 63]), ("mandelbulb", 96, 5.0, [0, 48, 95])])
def test_z_slabs_reassemble(ctx, name, res, bounds, splits):
    """z-slab decomposition (what N ranks do): per-slab results with halo recompute + global
    vertex offsets concatenate to exactly the single-slab result"""
    p, _ = s2m.params_from_cli(res, bounds)
    full = s2m.mesh_run(ctx, module_for(ctx, name), p)
    df = full.data()
    keys, quads, pos, ninv = [], [], [], 0
    base = 0
    for zb, ze in zip(splits[:-1], splits[1:]):
        p.z_begin, p.z_end = zb, ze
        r = s2m.mesh_begin(ctx, module_for(ctx, name), p)
        n_own = r.info().n_vertices
        r.finish(base)
        d = r.data()
        keys.append(d.keys.copy()); quads.append(d.quads.copy()); pos.append(d.positions.copy())
        ninv += d.n_invalid_quads
        base += n_own
        r.free()
    assert np.array_equal(np.concatenate(keys), df.keys)
    assert np.array_equal(np.concatenate(quads), df.quads)
    assert f32_equal(np.concatenate(pos), df.positions).all()
    assert ninv == df.n_invalid_quads
    full.free()


def test_bigger_sizes_properties(ctx):
    """size-independent properties at a size the oracle would not finish quickly: sorted unique keys,
    quads reference existing vertices, every quad's 4 cells are the right neighbours, determinism"""
    p, _ = s2m.params_from_cli(512, 5.0)
    m = module_for(ctx, "mandelbulb")
    r = s2m.mesh_run(ctx, m, p)
    d = r.data()
    k = d.keys
    assert len(k) > 400000
    assert np.all(k[1:] > k[:-1])
    q = d.quads
    assert q.max() < len(k)
    x = (k & 0xFFFF).astype(np.int64); y = ((k >> 16) & 0xFFFF).astype(np.int64); z = (k >> 32).astype(np.int64)
    cx, cy, cz = x[q], y[q], z[q]
    # the four cells of a quad span a 2x2x1 block
    span = (cx.max(1) - cx.min(1)) + (cy.max(1) - cy.min(1)) + (cz.max(1) - cz.min(1))
    assert np.all(span == 2)
    r2 = s2m.mesh_run(ctx, m, p)
    d2 = r2.data()
    assert np.array_equal(d2.keys, k) and np.array_equal(d2.quads, q) and f32_equal(d2.positions, d.positions).all()
    key = "mandelbulb_r512_b5_f0"
    if key in GOLDEN:
        assert mesh_digests(d.positions, d.normals, d.keys, d.nibbles, d.quads, d.n_invalid_quads) == GOLDEN[key]
    r.free()
    r2.free()


@pytest.mark.parametrize("name,dims,bmin,bmax", [
    ("torus", (70, 45, 33), [-1.0, -0.6, -0.5], [1.0, 0.7, 0.5]),
    ("mandelbulb", (96, 96, 96), [-2.5, -2.5, -2.5], [2.5, 2.5, 2.5]),
    ("p_key", (64, 64, 64), [-1, -1, -1], [1, 1, 1]),
    ("torus", (31, 33, 2), [-1, -1, -0.1], [1, 1, 0.1]),
])
def test_classify_paths_agree(ctx, name, dims, bmin, bmax):
    """K2 from K1's class bit planes (default) == K2 re-reading the f32 slab through shared memory"""
    out = []
    for flags in (s2m.MESH_KEEP_CANDIDATES, s2m.MESH_KEEP_CANDIDATES | s2m.MESH_CLASSIFY_FROM_SLAB):
        for budget in (0, (dims[0] + 32) * (dims[1] + 1) * 4 * 5):
            p = s2m.make_params(dims, bmin, bmax, flags=flags, slab_budget_bytes=budget)
            r = s2m.mesh_run(ctx, module_for(ctx, name), p)
            d = r.data()
            out.append((d.candidates.copy(), d.keys.copy(), d.quads.copy()))
            r.free()
    for o in out[1:]:
        assert np.array_equal(o[0], out[0][0]) and np.array_equal(o[1], out[0][1]) and np.array_equal(o[2], out[0][2])


def test_4096_cubed_properties(ctx):
    """BASELINE config 5: 4096^3 (beyond the reference's 2048 texture limit) on ONE GPU, chunked slab.
    No oracle at this size (~20 CPU-minutes): structural properties + determinism of the counts."""
    p, _ = s2m.params_from_cli(4096, 5.0, flags=s2m.MESH_NO_NORMALS)
    m = module_for(ctx, "mandelbulb")
    r = s2m.mesh_run(ctx, m, p)
    d = r.data()
    k = d.keys
    assert len(k) > 30_000_000 and d.timings["chunks"] > 8
    assert np.all(k[1:] > k[:-1])
    z = (k >> np.uint64(32)).astype(np.int64)
    assert z.min() >= 1 and z.max() <= 4095
    q = d.quads
    assert int(q.max()) < len(k)
    sel = np.random.default_rng(0).integers(0, len(q), 2_000_000)
    qs = q[sel]
    x = (k & np.uint64(0xFFFF)).astype(np.int64); y = ((k >> np.uint64(16)) & np.uint64(0xFFFF)).astype(np.int64)
    cx, cy, cz = x[qs], y[qs], z[qs]
    span = (cx.max(1) - cx.min(1)) + (cy.max(1) - cy.min(1)) + (cz.max(1) - cz.min(1))
    assert np.all(span == 2)
    counts = (len(k), len(q), d.n_invalid_quads)
    r.free()
    r = s2m.mesh_run(ctx, m, p)
    d = r.data()
    assert (len(d.keys), len(d.quads), d.n_invalid_quads) == counts
    r.free()


SLAB_GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "slab_digests.json")))


@pytest.mark.parametrize("key", sorted(SLAB_GOLDEN))
def test_sampled_slabs_match_oracle_digests(ctx, key):
    """BASELINE config 5 (mandelmesh.frag at 4096^3, one GPU, chunked slab, 64-bit indices), config 4 and config 3
    as literally stated (p_key --resolution 1024, default --bounds 2): the WHOLE grid is meshed on the GPU, then
    z-slab extracts (vertices, and quads as cell-key tuples) are compared with digests of oracle.mesh_run(z, z+2)
    on the same slabs (tests/golden/make_slab_golden.py).  At 4096^3 the oracle cannot be run in full inside a
    test (~20 CPU-minutes); 32 evenly spaced slab pairs pin the arithmetic, the ordering and the connectivity."""
    from tests.support.slabs import extract_slab, slab_digest
    name, r_, b_, _f = key.rsplit("_", 3)
    res, bounds = int(r_[1:]), float(b_[1:])
    p, _ = s2m.params_from_cli(res, bounds)
    r = s2m.mesh_run(ctx, module_for(ctx, name), p)
    d = r.data()
    assert np.all(d.keys[1:] > d.keys[:-1])
    qmax = d.quads.max(axis=1) if len(d.quads) else np.zeros(0, np.uint64)   # the emitting vertex: non-decreasing
    assert np.all(qmax[1:] >= qmax[:-1])
    bad = []
    n_checked = 0
    for z, want in sorted(SLAB_GOLDEN[key]["slabs"].items(), key=lambda kv: int(kv[0])):
        z = int(z)
        i0, i1 = extract_slab(d.keys, 1, z, z + SLAB_GOLDEN[key]["slices_per_slab"])
        j0, j1 = int(np.searchsorted(qmax, i0)), int(np.searchsorted(qmax, i1))
        got = slab_digest(d.keys[i0:i1], d.nibbles[i0:i1], d.positions[i0:i1], d.normals[i0:i1], d.quads[j0:j1], index_base=i0)
        n_checked += got["n_vertices"]
        if got != want:
            bad.append((z, got, want))
    r.free()
    assert not bad, bad[:3]
    assert n_checked > 0


@pytest.mark.parametrize("name,res,bounds", [("torus", 64, 1.0), ("mandelbulb", 128, 2.0)])
def test_invalid_quad_records(ctx, name, res, bounds):
    """f4: which quads the reference would report as invalid (mesh.rs:270-278), in its order"""
    p, _ = s2m.params_from_cli(res, bounds, flags=s2m.MESH_KEEP_INVALID)
    r = s2m.mesh_run(ctx, module_for(ctx, name), p)
    o = oracle.mesh_run(name, res, bounds)
    d = r.data()
    want = o.invalid_records()
    assert d.n_invalid_quads == o.n_invalid_quads == len(want) and len(want) > 0
    assert np.array_equal(d.invalid_records, want)
    assert_same(d, o, f"{name} with invalid quads")
    r.free()
    o.free()


@pytest.mark.parametrize("name,res,bounds", [("torus", 64, 1.0), ("mandelbulb", 128, 5.0), ("p_key", 64, 20.0), ("martin_cube", 64, 2.0)])
def test_consistent_corner_mode_matches_oracle(ctx, name, res, bounds):
    """S2M_MESH_CONSISTENT_CORNERS (with ALL_SLICES: the CLI's --watertight): cell max corner = next
    cell's min corner.  Not the reference's arithmetic, so the oracle carries the same switch."""
    flags = s2m.MESH_ALL_SLICES | s2m.MESH_CONSISTENT_CORNERS
    p, _ = s2m.params_from_cli(res, bounds, flags=flags)
    r = s2m.mesh_run(ctx, module_for(ctx, name), p)
    o = oracle.mesh_run(name, res, bounds, flags=oracle.FLAG_ALL_SLICES | oracle.FLAG_CONSISTENT_CORNERS)
    try:
        assert_same(r.data(), o, f"{name} {res}^3 consistent corners")
    finally:
        r.free()
        o.free()


def test_watertight_mode_closes_the_2048_mesh_cracks(ctx):
    """at 1024^3 the faithful mandelbulb mesh loses quads to 1-ulp corner disagreements (F4) and the
    slice lag (F3); the watertight mode loses none, and every mesh edge is then shared by an even
    number of quads (closed surface inside the box)"""
    m = module_for(ctx, "mandelbulb")
    p, _ = s2m.params_from_cli(1024, 5.0, flags=s2m.MESH_ALL_SLICES | s2m.MESH_CONSISTENT_CORNERS | s2m.MESH_NO_NORMALS)
    r = s2m.mesh_run(ctx, m, p)
    d = r.data()
    assert d.n_invalid_quads == 0 and len(d.quads) > 2_000_000
    q = d.quads.astype(np.int64)
    e = np.concatenate([np.stack([q[:, i], q[:, (i + 1) % 4]], 1) for i in range(4)])
    e.sort(axis=1)
    code = e[:, 0] * (len(d.keys) + 1) + e[:, 1]
    _, counts = np.unique(code, return_counts=True)
    assert (counts % 2 == 0).all() and (counts == 2).mean() > 0.99  # one vertex per cell: a few edges are shared by 4
    r.free()


def test_u32_quad_indices(ctx, tmp_path):
    """S2M_MESH_QUADS_U32: the reference's own index type (lib.rs Quad(u32, ..)); same values, same files,
    and an index that would not fit is refused"""
    name, res, bounds = "mandelbulb", 128, 5.0
    m = module_for(ctx, name)
    p, _ = s2m.params_from_cli(res, bounds, flags=s2m.MESH_QUADS_U32)
    r = s2m.mesh_run(ctx, m, p)
    o = oracle.mesh_run(name, res, bounds)
    d = r.data()
    assert d.quads.dtype == np.uint32
    assert_same(d, o, "u32 quads")
    ref = tmp_path / "ref.ply"
    o.write_ply(ref)
    r.write_mesh(tmp_path / "one.ply")
    assert (tmp_path / "one.ply").read_bytes() == ref.read_bytes()
    # z-slabs with u32 indices and global bases
    parts, base = [], 0
    for zb, ze in ((0, 50), (50, 90), (90, 127)):
        p.z_begin, p.z_end = zb, ze
        s = s2m.mesh_begin(ctx, m, p)
        n_own = s.info().n_vertices
        s.finish(base)
        base += n_own
        parts.append(s)
    assert np.array_equal(np.concatenate([s.data().quads for s in parts]), o.quads)
    s2m.write_mesh_parts(parts, tmp_path / "parts.ply")
    assert (tmp_path / "parts.ply").read_bytes() == ref.read_bytes()
    for s in parts:
        s.free()
    p.z_begin, p.z_end = 0, 50
    s = s2m.mesh_begin(ctx, m, p)
    with pytest.raises(s2m.S2mError) as e:
        s.finish(2 ** 32 - 5)
    assert "32 bits" in str(e.value)
    s.free()
    r.free()
    o.free()


@pytest.mark.parametrize("seed", range(12))
def test_randomized_configurations(ctx, seed):
    """random grid shapes, boxes, flag combinations, slab budgets (1 .. many z-chunks, so both the single-
    stream and the two-stream chunk pipeline) and z-slab splits, each against the oracle; every
    configuration runs twice on the same context (the second run streams its output copies)"""
    rng = np.random.default_rng(1000 + seed)
    name = ["torus", "mandelbulb", "martin_cube", "p_key"][seed % 4]
    scale = {"torus": 1.0, "mandelbulb": 1.6, "martin_cube": 1.2, "p_key": 9.0}[name]
    dims = tuple(int(v) for v in rng.integers(9, 72, 3))
    centre = rng.uniform(-0.15, 0.15, 3) * scale
    half = rng.uniform(0.7, 1.3, 3) * scale
    bmin, bmax = (centre - half).astype(np.float32), (centre + half).astype(np.float32)
    all_slices = bool(rng.integers(0, 2))
    consistent = bool(rng.integers(0, 2))
    flags = (s2m.MESH_ALL_SLICES if all_slices else 0) | (s2m.MESH_CONSISTENT_CORNERS if consistent else 0)
    flags |= s2m.MESH_QUADS_U32 if rng.integers(0, 2) else 0
    flags |= s2m.MESH_CLASSIFY_FROM_SLAB if rng.integers(0, 4) == 0 else 0
    oflags = (oracle.FLAG_ALL_SLICES if all_slices else 0) | (oracle.FLAG_CONSISTENT_CORNERS if consistent else 0)
    plane_bytes = ((dims[0] + 1 + 31) // 32 * 32) * (dims[1] + 1) * 4
    budget = int(plane_bytes * rng.choice([2, 3, 5, 9, 1000])) + 64
    n_slices = dims[2] if all_slices else dims[2] - 1
    cuts = sorted(set([0, n_slices] + [int(c) for c in rng.integers(1, max(2, n_slices), rng.integers(0, 3))]))
    o = oracle.mesh_run(name, dims, flags=oflags, bmin=bmin, bmax=bmax)
    m = module_for(ctx, name)
    try:
        for rep in range(2):
            p = s2m.make_params(dims, bmin, bmax, flags=flags, slab_budget_bytes=budget)
            r = s2m.mesh_run(ctx, m, p)
            assert_same(r.data(), o, f"seed {seed} rep {rep}: {name} {dims} flags {flags} budget {budget}")
            r.free()
        parts, base = [], 0
        for zb, ze in zip(cuts[:-1], cuts[1:]):
            p = s2m.make_params(dims, bmin, bmax, flags=flags, slab_budget_bytes=budget, z_begin=zb, z_end=ze)
            s = s2m.mesh_begin(ctx, m, p)
            n_own = s.info().n_vertices
            s.finish(base)
            base += n_own
            parts.append(s)
        what = f"seed {seed}: {name} {dims} cuts {cuts} flags {flags}"
        if parts:
            assert np.array_equal(np.concatenate([s.data().keys for s in parts]), o.keys), what
            assert np.array_equal(np.concatenate([s.data().quads for s in parts]).astype(np.uint64).reshape(-1, 4), o.quads), what
            assert f32_equal(np.concatenate([s.data().positions for s in parts]), o.positions).all(), what
            assert sum(s.data().n_invalid_quads for s in parts) == o.n_invalid_quads, what
        for s in parts:
            s.free()
    finally:
        o.free()


def test_kernel_timing_spans_are_opt_in(ctx):
    """S2M_MESH_TIMINGS: the per-kernel fields of s2m_timings come from CUDA-event spans around every launch; off by default
    (device_ms / total_ms / host_wall_ms / launches / chunks are always there).  Same mesh either way."""
    mod = module_for(ctx, "mandelbulb")
    p, _ = s2m.params_from_cli(256, 5.0)
    r = s2m.mesh_run(ctx, mod, p)
    d0 = r.data()
    t = d0.timings
    assert t["k1_slab_ms"] == 0 and t["k4_vertices_ms"] == 0 and t["d2h_ms"] == 0
    assert t["device_ms"] > 0 and t["total_ms"] >= t["device_ms"] and t["host_wall_ms"] > 0 and t["launches"] == 5 * t["chunks"]
    keys0, quads0 = d0.keys.copy(), d0.quads.copy()
    r.free()
    p.flags |= s2m.MESH_TIMINGS
    r = s2m.mesh_run(ctx, mod, p)
    d1 = r.data()
    t = d1.timings
    assert min(t[k] for k in ("k1_slab_ms", "k2_classify_ms", "k3_compact_ms", "k4_vertices_ms", "k4_quads_ms", "d2h_ms")) > 0
    assert np.array_equal(d1.keys, keys0) and np.array_equal(d1.quads, quads0)
    r.free()
