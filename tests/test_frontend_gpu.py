"""GPU tests of the front-end: the CUDA C++ it emits gives the same bits when NVRTC compiles it for
sm_100a (--fmad=false) as when g++ compiles the same text for the host (-ffp-contract=off), for
programs that use every language feature the CPU-side tests cover; and the reference's own test
shader (src/lib.rs test_naga) meshes to its golden digest."""
import json
import os
import textwrap

import numpy as np
import pytest

import sdf2mesh_b200 as s2m
from tests.conftest import ROOT
from tests.support import host_eval
from tests.support.digest import f32_equal, mesh_digests
from tests.test_frontend import NAGA_TEST_GLSL

pytestmark = pytest.mark.gpu

WGSL_PROGRAMS = {
    # NaN in a region of space: a NaN corner is neither "outside" (v > tau) nor "inside" (v < -tau) for the classifier, and
    # "inside" (!(v > 0)) for the reference's cell rule
    "nan_region": """
        fn sdf3d(p: vec3f) -> f32 {
          let d = length(p) - 0.8;
          let hole = sqrt(0.09 - (p.x - 0.5) * (p.x - 0.5) - p.y * p.y);
          return select(d, d + hole * 0.0, p.z > 0.2);
        }""",
    "control_flow": """
        fn fold(q: vec3f) -> vec3f { var v = abs(q); if (v.x < v.y) { v = v.yxz; } if (v.x < v.z) { v = v.zyx; } return v; }
        fn sdf3d(p: vec3f) -> f32 {
          var z = p; var acc = 0.0; var i = 0;
          loop { if (i >= 6) { break; } z = fold(z) * 1.7 - vec3f(0.6, 0.2, 0.1); acc += sin(z.x * 3.0) * cos(z.x * 3.0);
                 continuing { i++; break if length(z) > 20.0; } }
          while (acc > 2.0) { acc = acc - 1.5; }
          return length(z) * pow(1.7, -f32(i)) - 0.05 + 0.01 * acc;
        }""",
    "math_mix": """
        fn sdf3d(p: vec3f) -> f32 {
          let a = atan2(p.y, p.x); let r = length(p.xy);
          let q = vec3f(r * cos(a * 3.0), r * sin(a * 3.0), p.z);
          let e = exp(-abs(q.z)) + log(1.0 + r) + tanh(q.x) + sqrt(abs(q.y)) + asin(clamp(q.z * 0.3, -1.0, 1.0)) + acos(clamp(q.x * 0.2, -1.0, 1.0));
          return smoothstep(0.0, 4.0, e) + fract(q.x) * 0.1 + mix(q.y, q.z, 0.25) * 0.01 + sign(q.x) * step(0.5, r) * 0.001 + pow(r + 0.5, 2.5) * 1e-3 + exp2(-r) + log2(r + 1.0);
        }""",
    "aggregates": """
        const N = 3;
        struct Hit { d: f32, id: i32, }
        struct Scene { spheres: array<vec4f, N>, best: Hit, }
        const RADII = array<f32, N>(0.5, 0.25, 0.75);
        var<private> evals: i32;
        fn closer(a: Hit, b: Hit) -> Hit { evals = evals + 1; if (a.d < b.d) { return a; } return b; }
        fn sdf3d(p: vec3f) -> f32 {
          var sc: Scene;
          for (var i = 0; i < N; i++) { sc.spheres[i] = vec4f(f32(i) - 1.0, 0.0, 0.0, RADII[i]); }
          sc.best = Hit(1e9, -1);
          for (var i = 0; i < N; i++) { let sp = sc.spheres[i]; sc.best = closer(sc.best, Hit(length(p - sp.xyz) - sp.w, i)); }
          var v = p;
          v[sc.best.id] = v[sc.best.id] * 2.0;
          var w = 0.0;
          switch sc.best.id { case 0: { w = 1.0; } case 1, 2: { w = 2.0; } default: { w = -1.0; } }
          let m = mat3x3f(v, p, v + p);
          return sc.best.d + 0.001 * w + 0.01 * m[sc.best.id].y + 0.0001 * f32(evals);
        }""",
}


GLSL_PROGRAMS = {
    "integer_hash_noise": """#version 450 core
        uvec3 pcg3d(uvec3 v) {
          v = v * 1664525u + 1013904223u;
          v.x += v.y * v.z; v.y += v.z * v.x; v.z += v.x * v.y;
          v ^= v >> 16u;
          v.x += v.y * v.z; v.y += v.z * v.x; v.z += v.x * v.y;
          return v;
        }
        vec3 hash33(vec3 p) { uvec3 q = pcg3d(uvec3(ivec3(floor(p)) + 1000)); return vec3(q) * (1.0 / float(0xffffffffu)); }
        float vnoise(vec3 p) {
          vec3 i = floor(p), f = fract(p);
          vec3 u = f * f * (3.0 - 2.0 * f);
          float acc = 0.0;
          for (int k = 0; k < 8; ++k) {
            ivec3 o = ivec3(k & 1, (k >> 1) & 1, (k >> 2) & 1);
            vec3 w = mix(1.0 - u, u, equal(o, ivec3(1)));
            acc += hash33(i + vec3(o)).x * w.x * w.y * w.z;
          }
          return acc;
        }
        float sdf(vec3 p) {
          bvec3 neg = lessThan(p, vec3(0.0));
          vec3 q = mix(p, -p, neg);
          int m = int(floatBitsToUint(q.x) >> 23u) & 0xff;
          ivec3 c = clamp(ivec3(p * 2.0), ivec3(-2), ivec3(2));
          return length(q) - 1.0 + 0.05 * vnoise(p * 4.0) + float(m - 127) * 0.001 + 0.001 * float(abs(c.x) + max(c.y, c.z) + sign(c.z) + 7 / (c.x - c.x))
                 + refract(normalize(p + 0.1), vec3(0.0, 1.0, 0.0), 0.9).x * 1e-3;
        }
        void main() {}
        """,
}


@pytest.mark.parametrize("name", sorted(WGSL_PROGRAMS) + sorted(GLSL_PROGRAMS))
def test_device_bits_equal_host_bits(ctx, name, tmp_path):
    if name in GLSL_PROGRAMS:
        frag = tmp_path / (name + ".frag")
        frag.write_text(textwrap.dedent(GLSL_PROGRAMS[name]))
        sh = s2m.Sdf3DShader.from_glsl_fragment_shader(frag, "sdf")
    else:
        sh = s2m.Sdf3DShader.from_source(textwrap.dedent(WGSL_PROGRAMS[name]))
    cuda = sh.lower_to_cuda()
    rng = np.random.default_rng(11)
    pts = rng.uniform(-3, 3, (200_000, 3)).astype(np.float32)
    dev = sh.create_shader_module(ctx).eval_points(pts)
    host = host_eval.eval_points(cuda, pts)
    eq = f32_equal(dev, host)
    assert eq.all(), f"{np.count_nonzero(~eq)} of {len(pts)} values differ"


def test_naga_test_shader_meshes_to_its_golden_digest(ctx, tmp_path):
    """the GLSL of the reference's only unit test (lib.rs test_naga), through GLSL -> WGSL -> CUDA -> mesh"""
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", "digests.json")))
    keys = [k for k in golden if k.startswith("naga_sphere_")]
    assert keys
    w = s2m.WgslShaderCode(s2m.convert_glsl_to_wgsl(NAGA_TEST_GLSL))
    w.remove_function("fn main_1(")
    w.remove_function("fn main(")
    w.remove_line("@fragment")
    m = s2m.Sdf3DShader.from_source(w.text, s2m.SRC_WGSL).create_shader_module(ctx)
    for key in keys:
        _, r_, b_, f_ = key.rsplit("_", 3)
        p, _ = s2m.params_from_cli(int(r_[1:]), float(b_[1:]), flags=s2m.MESH_ALL_SLICES if int(f_[1:]) & 1 else 0)
        r = s2m.mesh_run(ctx, m, p)
        d = r.data()
        assert mesh_digests(d.positions, d.normals, d.keys, d.nibbles, d.quads, d.n_invalid_quads) == golden[key]
        r.free()


def _program(name, tmp_path):
    if name.startswith("shadertoy_"):
        return s2m.Sdf3DShader.from_shadertoy_source(open(os.path.join(ROOT, "tests", "data", name + ".glsl")).read(), "map")
    if name in GLSL_PROGRAMS:
        frag = tmp_path / (name + ".frag")
        frag.write_text(textwrap.dedent(GLSL_PROGRAMS[name]))
        return s2m.Sdf3DShader.from_glsl_fragment_shader(frag, "sdf")
    return s2m.Sdf3DShader.from_source(textwrap.dedent(WGSL_PROGRAMS[name]))


@pytest.mark.parametrize("name,res,bounds", [("aggregates", 48, 4.0), ("control_flow", 40, 3.0), ("integer_hash_noise", 48, 3.0), ("nan_region", 48, 2.5),
                                             ("shadertoy_raymarcher", 64, 3.0), ("shadertoy_idioms", 64, 3.5)])
def test_generated_programs_mesh_like_the_oracle(ctx, tmp_path, name, res, bounds):
    """whole-path parity for shaders without a hand transcription: the oracle runs its cell loop on the
    front-end's emitted code compiled for the host (SDF id "plugin"), the GPU path on the same code
    compiled by NVRTC; keys, nibbles, quads, positions and normals must agree bit for bit"""
    import oracle
    from tests.test_parity_gpu import assert_same
    sh = _program(name, tmp_path)
    oracle.set_plugin(host_eval.scalar_function(sh.lower_to_cuda()))
    try:
        for flags, oflags in ((0, 0), (s2m.MESH_ALL_SLICES | s2m.MESH_CONSISTENT_CORNERS, oracle.FLAG_ALL_SLICES | oracle.FLAG_CONSISTENT_CORNERS)):
            o = oracle.mesh_run("plugin", res, bounds, flags=oflags)
            bmin, bmax = oracle.cube_bounds(bounds)
            p = s2m.make_params(res, bmin, bmax, flags=flags)   # not params_from_cli: that rounds res up to a power of two
            r = s2m.mesh_run(ctx, sh.create_shader_module(ctx), p)
            assert len(o.keys) > 300, "the test box should contain some surface"
            assert_same(r.data(), o, f"{name} flags {flags}")
            r.free()
            o.free()
    finally:
        oracle.set_plugin(None)
