"""GPU tests of the packed f32x2 form of K1 (csrc/s2m_pvec.h: two grid corners per evaluation, FFMA2 /
FMUL2 on sm_100a).  Bar: the packed form gives, lane for lane, the BITS of the scalar form -- on the
device (FFMA2 arithmetic) exactly as in the host emulation the CPU suite checks -- and a mesh made
with a packed K1 is the oracle's mesh.

mandelmesh.frag uses the packed K1 by default (7 transcendental calls), so every mandelbulb case of
tests/test_parity_gpu.py -- up to the 2048^3 golden digest -- runs through it as well; here the packed
form is forced for the SDFs that would not get it by default."""
import os
import textwrap

import numpy as np
import pytest

import oracle
import sdf2mesh_b200 as s2m
from tests.conftest import load_example_shader
from tests.support import host_eval
from tests.support.digest import f32_equal
from tests.test_frontend_gpu import GLSL_PROGRAMS, WGSL_PROGRAMS
from tests.test_parity_gpu import assert_same

pytestmark = pytest.mark.gpu

# "1": sqrt per lane, "2": sqrt's refinement step in f32x2 as well (S2M_TEST_PACKED_VARIANT selects)
VARIANT = os.environ.get("S2M_TEST_PACKED_VARIANT", "2")

PROGRAMS = ["torus", "martin_cube", "p_key", "mandelbulb", "wgsl:control_flow", "wgsl:math_mix", "glsl:integer_hash_noise"]


def shader_for(name, tmp_path):
    if name.startswith("wgsl:"):
        return s2m.Sdf3DShader.from_source(textwrap.dedent(WGSL_PROGRAMS[name[5:]]))
    if name.startswith("glsl:"):
        frag = tmp_path / "p.frag"
        frag.write_text(textwrap.dedent(GLSL_PROGRAMS[name[5:]]))
        return s2m.Sdf3DShader.from_glsl_fragment_shader(frag, "sdf")
    return load_example_shader(name)


@pytest.mark.parametrize("name", PROGRAMS)
def test_packed_device_bits(ctx, name, tmp_path, monkeypatch):
    """lane lo == scalar value, always; lane hi == scalar value unless the lanes disagreed; the same
    pairs disagree on the device as in the host emulation"""
    monkeypatch.setenv("S2M_K1_PACKED", VARIANT)
    sh = shader_for(name, tmp_path)
    mod = sh.create_shader_module(ctx)
    assert mod.packed, mod.log
    rng = np.random.default_rng(5)
    n = 150_000
    a = rng.uniform(-2.5, 2.5, (n, 3)).astype(np.float32)
    near = a.copy()
    near[:, 0] += np.float32(5.0 / 2047.0)
    far = np.roll(a, 1, axis=0)
    want_a = mod.eval_points(a)
    packed_text = sh.lower_to_cuda_packed()
    for b in (near, far):
        want_b = mod.eval_points(b)
        lo, hi, dv = mod.eval_pairs(a, b)
        eq = f32_equal(lo, want_a)
        assert eq.all(), f"lane lo: {np.count_nonzero(~eq)} of {n} values differ from the scalar kernel"
        eq = f32_equal(hi, want_b) | dv
        assert eq.all(), f"lane hi: {np.count_nonzero(~eq)} of {n} values differ although the lanes agreed"
        h_lo, h_hi, h_dv = host_eval.eval_pairs(packed_text, a, b)
        assert np.array_equal(dv, h_dv), "device and host emulation disagree on which pairs diverge"
        assert f32_equal(lo, h_lo).all() and (f32_equal(hi, h_hi) | dv).all()


def test_packed_division_by_one(ctx, monkeypatch):
    """s2m_pvec.h p_div: a divisor pair of exactly (1, 1) takes one packed multiplication by 1 instead of two IEEE
    divisions.  Same bits as the scalar quotient for every numerator -- -0, denormals, infinities, NaN (canonical either
    way), huge and tiny values -- and for divisor pairs of which only one lane is 1"""
    monkeypatch.setenv("S2M_K1_PACKED", VARIANT)
    src = ("fn sdf3d(p: vec3f) -> f32 { var d = 1.0; if (p.z > 0.5) { d = p.z; } let num = p.x * p.y; let a = num / d; "
           "let b = (p.x / p.y) / d; let c = (num - num) / (p.y - p.y) / d; if (p.z == 0.0) { return a; } if (p.z == 0.25) { return b; } "
           "if (p.z == 0.5) { return c; } return a + b * 0.5; }")
    mod = s2m.Sdf3DShader.from_source(src).create_shader_module(ctx)
    assert mod.packed, mod.log
    special = np.array([0.0, -0.0, 1.0, -1.0, 1e-45, -1e-45, 1e-38, 3e38, -3e38, np.inf, -np.inf, np.nan, 1e-20, 1e20, 0.5, 2.0], np.float32)
    xs, ys, zs = np.meshgrid(special, special, np.array([0.0, 0.25, 0.5, 1.0, 1.0000001, 3.0, np.nan], np.float32), indexing="ij")
    a = np.stack([xs.ravel(), ys.ravel(), zs.ravel()], 1).astype(np.float32)
    want_a = mod.eval_points(a)
    for shift in (7, 16 * 7, 3 * 16 * 7 + 5 * 7, 1, 3):   # z is the fastest index: multiples of 7 pair points with the same divisor
        b = np.roll(a, shift, axis=0)
        want_b = mod.eval_points(b)
        lo, hi, dv = mod.eval_pairs(a, b)
        assert f32_equal(lo, want_a).all()
        agree = dv == 0
        assert f32_equal(hi[agree], want_b[agree]).all()
        if shift % 7 == 0:
            assert agree.mean() > 0.8   # only NaN z compares differently from itself... it does not: same branch in both lanes
    assert np.isnan(want_a).any() and np.isinf(want_a).any() and (want_a == 0).any()


@pytest.mark.parametrize("name,res,bounds", [("torus", 128, 2.0), ("martin_cube", 128, 2.0), ("p_key", 128, 20.0), ("mandelbulb", 128, 5.0)])
def test_mesh_with_packed_k1_matches_oracle(ctx, name, res, bounds, monkeypatch):
    monkeypatch.setenv("S2M_K1_PACKED", VARIANT)
    mod = load_example_shader(name).create_shader_module(ctx)
    assert mod.packed
    p, _ = s2m.params_from_cli(res, bounds)
    r = s2m.mesh_run(ctx, mod, p)
    o = oracle.mesh_run(name, res, bounds)
    try:
        assert_same(r.data(), o, f"{name} {res}^3, packed K1")
    finally:
        r.free()
        o.free()


@pytest.mark.parametrize("name,bounds", [("torus", 2.0), ("mandelbulb", 5.0), ("p_key", 20.0)])
def test_packed_slab_equals_scalar_slab(ctx, name, bounds, monkeypatch):
    """K1's slab, plane by plane: packed kernel (pairs + re-evaluated corners) == scalar kernel"""
    res = 200  # not a multiple of the tile: padded pitch, partial tiles
    bmin, bmax = oracle.cube_bounds(bounds)
    p = s2m.make_params(res, bmin, bmax)
    monkeypatch.setenv("S2M_K1_PACKED", "0")
    scalar = load_example_shader(name).create_shader_module(ctx)
    monkeypatch.setenv("S2M_K1_PACKED", VARIANT)
    packed = load_example_shader(name).create_shader_module(ctx)
    assert packed.packed and not scalar.packed
    for plane in (0, 61, 100, res):
        a = s2m.debug_slab_plane(ctx, scalar, p, plane)
        b = s2m.debug_slab_plane(ctx, packed, p, plane)
        eq = f32_equal(a, b)
        assert eq.all(), f"plane {plane}: {np.count_nonzero(~eq)} corners differ"


def test_default_policy(ctx):
    """packed by default only where it pays: SDFs dominated by transcendental functions"""
    assert load_example_shader("mandelbulb").create_shader_module(ctx).packed
    assert load_example_shader("torus").create_shader_module(ctx).packed          # tiny: gains as well
    assert not load_example_shader("p_key").create_shader_module(ctx).packed
    assert not load_example_shader("martin_cube").create_shader_module(ctx).packed
