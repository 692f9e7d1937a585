"""Front-end tests (no GPU): .sdf3d preprocessing, WGSL / GLSL lowering, WGSL text munging.

Emitted CUDA C++ is compiled with g++ (tests/support/host_eval.py) and evaluated on the CPU; the
hand-transcribed SDFs in oracle/sdf_examples.h are the independent answer."""
import os
import textwrap

import numpy as np
import pytest

import oracle
import sdf2mesh_b200 as s2m
from tests.conftest import EXAMPLES, ROOT, load_example_shader
from tests.support import host_eval
from tests.support.digest import f32_equal

REF_EXAMPLES = "/root/reference/examples"
RNG = np.random.default_rng(3)


def points(bounds, n=100_000):
    return RNG.uniform(-bounds / 2, bounds / 2, (n, 3)).astype(np.float32)


@pytest.mark.parametrize("name,bounds", [("torus", 2.0), ("martin_cube", 2.5), ("p_key", 20.0), ("mandelbulb", 5.0)])
def test_examples_lower_bit_exact(built, name, bounds):
    cuda = load_example_shader(name).lower_to_cuda()
    pts = points(bounds)
    assert f32_equal(host_eval.eval_points(cuda, pts), oracle.eval_points(name, pts)).all()


@pytest.mark.skipif(not os.path.isdir(REF_EXAMPLES), reason="reference checkout not present")
@pytest.mark.parametrize("name,f,bounds", [("torus", "torus.sdf3d", 2.0), ("martin_cube", "martin_cube.sdf3d", 2.5),
                                           ("p_key", "p_key.sdf3d", 20.0), ("mandelbulb", "mandelmesh.frag", 5.0)])
def test_our_examples_equal_reference_examples(built, name, f, bounds):
    """examples/ holds re-typed equivalents of the reference inputs; same values at every point"""
    path = os.path.join(REF_EXAMPLES, f)
    ref = s2m.Sdf3DShader.from_glsl_fragment_shader(path, "sdf") if f.endswith(".frag") else s2m.Sdf3DShader.from_path(path)
    pts = points(bounds, 50_000)
    a = host_eval.eval_points(ref.lower_to_cuda(), pts)
    b = host_eval.eval_points(load_example_shader(name).lower_to_cuda(), pts)
    assert f32_equal(a, b).all()


def _reference_expansion(path, src_dir):
    """Sdf3DShader::shader_source_input (shader.rs:159-203) restated: a `use X;` line becomes the bytes of the module
    files (module table :50-62), everything else is copied line by line with a newline"""
    mod = lambda f: open(os.path.join(src_dir, f), encoding="utf-8").read()
    table = {"sdf::*": ["sdf_op.wgsl"], "sdf::op": ["sdf_op.wgsl"], "sdf3d::normal": ["sdf3d_normal.wgsl"],
             "sdf3d::primitives": ["sdf3d_primitives.wgsl"], "sdf3d::*": ["sdf3d_primitives.wgsl", "sdf3d_normal.wgsl"]}
    out = ""
    for line in open(path, encoding="utf-8").read().splitlines():
        t = line.strip()
        if t.endswith(";") and t.startswith("use"):
            name = t.replace("use", "", 1).replace('"', "").replace(";", "").strip()
            out += "".join(mod(f) for f in table.get(name, []))
            continue
        assert not (t.endswith(";") and t.startswith("include"))   # the examples have none
        out += line + "\n"
    return out


@pytest.mark.skipif(not os.path.isdir(REF_EXAMPLES), reason="reference checkout not present")
@pytest.mark.parametrize("f,bounds", [("torus.sdf3d", 2.0), ("martin_cube.sdf3d", 2.5), ("p_key.sdf3d", 20.0)])
def test_reference_module_files_give_the_reference_dump(built, tmp_path, monkeypatch, f, bounds):
    """with S2M_WGSL_MODULE_DIR pointing at the reference's three library files, --debug-wgsl
    (Sdf3DShader::write_to_file, shader.rs:206-210) is byte for byte what the reference writes, the pasted library text is
    what gets compiled, and it evaluates to the same bits as the built-in device library"""
    path = os.path.join(REF_EXAMPLES, f)
    builtin = s2m.Sdf3DShader.from_path(path)
    monkeypatch.setenv("S2M_WGSL_MODULE_DIR", "/root/reference/src")
    sh = s2m.Sdf3DShader.from_path(path)
    dump = tmp_path / "dump.wgsl"
    sh.write_to_file(dump)
    assert dump.read_bytes() == _reference_expansion(path, "/root/reference/src").encode("utf-8")
    assert dump.read_bytes() != builtin.source.encode("utf-8")          # (the default dump carries this project's own library text)
    cuda = sh.lower_to_cuda()
    assert "u_sdf3d_normal" in cuda                                       # compiled from the pasted text, not linked from s2m_sdf3d_lib.h
    pts = points(bounds, 50_000)
    assert f32_equal(host_eval.eval_points(cuda, pts), host_eval.eval_points(builtin.lower_to_cuda(), pts)).all()
    monkeypatch.setenv("S2M_WGSL_MODULE_DIR", str(tmp_path / "nowhere"))  # files not found: the built-in text
    assert s2m.Sdf3DShader.from_path(path).source == builtin.source


@pytest.mark.skipif(not os.path.isdir(REF_EXAMPLES), reason="reference checkout not present")
def test_reference_normal_module_in_the_glsl_route(built, monkeypatch):
    """from_glsl_fragment_shader appends include_str!("sdf3d_normal.wgsl") with add_line (shader.rs:87)"""
    path = os.path.join(REF_EXAMPLES, "mandelmesh.frag")
    builtin = s2m.Sdf3DShader.from_glsl_fragment_shader(path, "sdf")
    monkeypatch.setenv("S2M_WGSL_MODULE_DIR", "/root/reference/src")
    sh = s2m.Sdf3DShader.from_glsl_fragment_shader(path, "sdf")
    normal = open("/root/reference/src/sdf3d_normal.wgsl", encoding="utf-8").read()
    assert (normal + "\n") in sh.source and sh.source.endswith("fn sdf3d(p: vec3<f32>) -> f32 { return sdf(p); }\n")
    pts = points(5.0, 20_000)
    assert f32_equal(host_eval.eval_points(sh.lower_to_cuda(), pts), host_eval.eval_points(builtin.lower_to_cuda(), pts)).all()


# --- the reference's own three unit tests (/root/reference/src/shadertoy.rs:354-453), restated -----
NAGA_WGSL = textwrap.dedent("""\
    fn mainImage(fragColor: ptr<function, vec4<f32>>, fragCoord: vec2<f32>) {
        var fragCoord_1: vec2<f32>;

        fragCoord_1 = fragCoord;
        return;
    }

    fn main_1() {
        return;
    }

    @fragment
    fn main() {
        main_1();
        return;
    }
    """)


def test_remove_function(built):
    w = s2m.WgslShaderCode("        \n" + NAGA_WGSL)
    w.remove_function("fn main_1()")
    assert "fn mainImage(fragColor" in w.text and "fn main_1()" not in w.text
    w.remove_function("fn main(")
    assert "fn mainImage(fragColor" in w.text and "fn main()" not in w.text
    w.remove_function("fn mainImage(")
    assert w.text.strip() == "@fragment"
    with pytest.raises(s2m.S2mError) as e:
        w.remove_function("fn nope(")
    assert e.value.kind == "SHADER" and "not found in shader" in str(e.value)


def test_rename_function(built):
    w = s2m.WgslShaderCode("fn normal(p_4: vec3<f32>, epsilon: f32) -> vec3<f32>")
    w.rename_function("normal", "sdf3d_normal")
    assert "fn sdf3d_normal(p_4: vec3<f32>, epsilon: f32) -> vec3<f32>" in w.text


NAGA_TEST_GLSL = "#version 450 core\n" + """
        layout(binding=0) uniform vec3      iResolution;           // viewport resolution (in pixels)
		layout(binding=0) uniform float     iTime;                 // shader playback time (in seconds)
		layout(binding=0) uniform float     iTimeDelta;            // render time (in seconds)
		layout(binding=0) uniform int       iFrame;                // shader playback frame
		layout(binding=0) uniform vec4      iChannelTime;          // channel playback time (in seconds)
		layout(binding=0) uniform vec4      iMouse;                // mouse pixel coords. xy: current (if MLB down), zw: click
		layout(binding=0) uniform vec4      iDate;                 // (year, month, day, time in seconds)
		layout(binding=0) uniform float     iSampleRate;           // sound sample rate (i.e., 44100)
""" + """
vec3 c = vec3(0.0, 0.0, 0.0);
const float r = 1.0;
float distance_from_sphere(vec3 p, vec3 c, float r)
{
    return distance(p, c) - r;
}

float sdf3d(vec3 p)
{
    float sphere_0 = distance_from_sphere(p, c, r);

    // set displacement
    float displacement = sin(5.0 * p.x) * sin(5.0 * p.y) * sin(5.0 * p.z) * 0.25 * sin(2.f * iTime);

    return sphere_0 + displacement;
}

vec3 sdf3d_normal(in vec3 p, in float epsilon)
{
    const vec3 small_step = vec3(epsilon, 0.0, 0.0);

    float gradient_x = sdf3d(p + small_step.xyy) - sdf3d(p - small_step.xyy);
    float gradient_y = sdf3d(p + small_step.yxy) - sdf3d(p - small_step.yxy);
    float gradient_z = sdf3d(p + small_step.yyx) - sdf3d(p - small_step.yyx);

    vec3 normal = vec3(gradient_x, gradient_y, gradient_z);

    return normalize(normal);
}

void mainImage( out vec4 fragColor, in vec2 fragCoord ) {}

""" + " void main() {}"


def test_naga_shader_converts_and_evaluates(built):
    """the reference's test only checks that conversion succeeds; here the result is also evaluated"""
    wgsl = s2m.convert_glsl_to_wgsl(NAGA_TEST_GLSL)
    assert "fn sdf3d(" in wgsl and "fn main_1(" in wgsl and "@fragment" in wgsl and "fn mainImage(" in wgsl
    w = s2m.WgslShaderCode(wgsl)
    w.remove_function("fn main_1(")
    w.remove_function("fn main(")
    w.remove_line("@fragment")
    cuda = s2m.Sdf3DShader.from_source(w.text, s2m.SRC_WGSL).lower_to_cuda()
    pts = points(2.5, 50_000)
    assert f32_equal(host_eval.eval_points(cuda, pts), oracle.eval_points("naga_sphere", pts)).all()


# --- .sdf3d directive semantics (/root/reference/src/shader.rs:159-203) ----------------------------
def test_use_directives(built, tmp_path):
    src = 'use sdf3d::*;\nuse nonexistent::module;\n  use "sdf::op";\nfn sdf3d(p: vec3f) -> f32 { return sdf3d_sphere(p, 1.0); }\n'
    sh = s2m.Sdf3DShader.from_source(src)
    assert "fn sdf3d_box(" in sh.source and "fn sdf3d_normal(" in sh.source and "fn sdf_op_smooth_union(" in sh.source
    assert "nonexistent" not in sh.source, "unknown modules are dropped silently (shader.rs:170-180)"
    assert sh.log.count("INFO ") == 3 and "INFO nonexistent::module" in sh.log
    assert sh.source.index("fn sdf3d_torus(") < sh.source.index("fn sdf3d_normal("), "sdf3d::* = primitives then normal"
    # only lines ending in ';' are directives; `used = 1.0;` is swallowed as a `use` line like in the reference
    sh2 = s2m.Sdf3DShader.from_source("use sdf3d::primitives\nused = 1.0;\nkeep me\n")
    assert sh2.source == "use sdf3d::primitives\nkeep me\n"


def test_include_directive(built, tmp_path):
    (tmp_path / "lib.wgsl").write_text("fn helper(x: f32) -> f32 { return x * 2.0; }\n")
    (tmp_path / "main.sdf3d").write_text('include "lib.wgsl";\ninclude "missing.wgsl";\nfn sdf3d(p: vec3f) -> f32 { return helper(p.x); }\n')
    cwd = os.getcwd()
    os.chdir(tmp_path)  # includes resolve against the CWD, not the including file (shader.rs:181-191)
    try:
        sh = s2m.Sdf3DShader.from_path("main.sdf3d")
    finally:
        os.chdir(cwd)
    assert "fn helper(" in sh.source and 'ERROR Could not include "missing.wgsl"' in sh.log
    vals = host_eval.eval_points(sh.lower_to_cuda(), np.array([[1.5, 0, 0]], np.float32))
    assert vals[0] == 3.0
    # unreadable top-level file: no exception, empty source, logged (shader.rs:44, :197-199)
    bad = s2m.Sdf3DShader.from_path(str(tmp_path / "nope.sdf3d"))
    assert bad.source == "" and "ERROR Could not include" in bad.log


def test_write_to_file_and_add_to_source(built, tmp_path):
    sh = s2m.Sdf3DShader.from_path(os.path.join(EXAMPLES, "torus.sdf3d"))
    before = sh.source
    sh.write_to_file(tmp_path / "debug.wgsl")  # --debug-wgsl (main.rs:223-225)
    assert (tmp_path / "debug.wgsl").read_text() == before
    sh.add_to_source("// tail\n")
    assert sh.source == before + "// tail\n"


# --- errors ------------------------------------------------------------------------------------
def test_glsl_missing_sdf(built, tmp_path):
    f = tmp_path / "x.frag"
    f.write_text("#version 450 core\nfloat other(vec3 p) { return p.x; }\nvoid main() {}\n")
    with pytest.raises(s2m.S2mError) as e:
        s2m.Sdf3DShader.from_glsl_fragment_shader(f, "sdf")
    assert e.value.kind == "MISSING_SDF"
    ok = s2m.Sdf3DShader.from_glsl_fragment_shader(f, "other")  # --glsl-sdf
    assert "fn sdf3d(p: vec3<f32>) -> f32 { return other(p); }" in ok.source
    f.write_text("#version 450 core\nfloat sdf(vec3 p) { return p.x; }\n")
    with pytest.raises(s2m.S2mError) as e:
        s2m.Sdf3DShader.from_glsl_fragment_shader(f, "sdf")  # no main(): naga rejects it too
    assert e.value.kind == "PARSE"


@pytest.mark.parametrize("src,kind", [
    ("fn sdf3d(p: vec3f) -> f32 { return p.x +; }", "PARSE"),
    ("fn sdf3d(p: vec3f) -> f32 { return q.x; }", "VALIDATION"),
    ("fn sdf3d(p: vec3f) -> f32 { return p; }", "VALIDATION"),
    ("fn sdf3d(p: vec3f) -> f32 { let a: i32 = 1.5; return p.x; }", "VALIDATION"),
    ("fn sdf3d(p: vec3f) -> f32 { return p.x + 1i; }", "VALIDATION"),
    ("fn other(p: vec3f) -> f32 { return p.x; }", "MISSING_SDF"),
    ("fn sdf3d(p: vec2f) -> f32 { return p.x; }", "MISSING_SDF"),
    ("fn sdf3d(p: vec3f) -> f32 { var t: texture_2d<f32>; return p.x; }", "UNSUPPORTED"),
    ("struct S { a: f32 }\nfn sdf3d(p: vec3f) -> f32 { let s = S(1.0); return s.b; }", "VALIDATION"),
    ("struct S { a: f32 }\nfn sdf3d(p: vec3f) -> f32 { let s = S(1.0, 2.0); return s.a; }", "VALIDATION"),
    ("struct S { a: S }\nfn sdf3d(p: vec3f) -> f32 { return p.x; }", "VALIDATION"),
    ("fn sdf3d(p: vec3f) -> f32 { let a = array<f32, 2>(1.0, 2.0); return a[2]; }", "VALIDATION"),
    ("fn sdf3d(p: vec3f) -> f32 { let a = array<f32, 2>(1.0, 2.0, 3.0); return a[0]; }", "VALIDATION"),
    ("fn sdf3d(p: vec3f) -> f32 { var a: array<f32>; return p.x; }", "UNSUPPORTED"),
    ("fn sdf3d(p: vec3f) -> f32 { switch 1 { case 1: { return p.x; } } return 0.0; }", "VALIDATION"),
    ("fn sdf3d(p: vec3f) -> f32 { switch 1 { case 1, 1: { return p.x; } default: { } } return 0.0; }", "VALIDATION"),
    ("fn sdf3d(p: vec3f) -> f32 { switch p.x { default: { } } return 0.0; }", "VALIDATION"),
])
def test_wgsl_errors(built, src, kind):
    with pytest.raises(s2m.S2mError) as e:
        s2m.Sdf3DShader.from_source(src).lower_to_cuda()
    assert e.value.kind == kind


def test_nvrtc_error_is_reported(built):
    sh = s2m.Sdf3DShader.from_source("float sdf3d(vec3 p) { return undefined_symbol(p); }", s2m.SRC_CUDA)
    with pytest.raises(s2m.S2mError) as e:
        sh.create_shader_module(None)
    assert e.value.kind == "NVRTC" and "undefined_symbol" in str(e.value)


# --- semantics -----------------------------------------------------------------------------------
def run_wgsl(src, pts):
    return host_eval.eval_points(s2m.Sdf3DShader.from_source(src).lower_to_cuda(), np.asarray(pts, np.float32))


def test_abstract_float_constants_fold_in_f64(built):
    """module consts are abstract floats: K*0.1 is folded in f64 and rounded once (naga), which differs
    from f32(K)*f32(0.45) for K = 3"""
    src = "const K = 3.0;\nfn sdf3d(p: vec3f) -> f32 { return p.x - K*0.45; }"
    v = run_wgsl(src, [[8.0, 0, 0]])[0]
    assert v == np.float32(8.0) - np.float32(3.0 * 0.45)
    assert np.float32(3.0 * 0.45) != np.float32(3.0) * np.float32(0.45)
    src2 = "const K: f32 = 3.0;\nfn sdf3d(p: vec3f) -> f32 { return p.x - K*0.45; }"  # concrete: f32 arithmetic
    assert run_wgsl(src2, [[8.0, 0, 0]])[0] == np.float32(8.0) - np.float32(3.0) * np.float32(0.45)


def test_wgsl_control_flow_and_swizzles(built):
    src = textwrap.dedent("""\
        fn fold(p: vec3f) -> vec3f {
            var q = p;
            for (var i = 0; i < 3; i++) {
                q = abs(q) - vec3f(0.5);
                if (q.x < q.y) { let t = q.x; q.x = q.y; q.y = t; }
                q.z *= 1.5;
            }
            return q;
        }
        fn sdf3d(p: vec3f) -> f32 {
            var acc = 0.0;
            var i: u32 = 0u;
            loop {
                if (i >= 4u) { break; }
                acc += f32(i) * 0.25;
                continuing { i = i + 1u; }
            }
            let q = fold(p).zyx;
            var k = 0;
            while (k < 2) { k += 1; if (k == 1) { continue; } acc -= 1.0; }
            return select(length(q.xy), max(q.x, q.z), p.y > 0.0) + acc + f32(k);
        }""")
    pts = points(3.0, 2000)

    def ref(p):
        q = p.astype(np.float32).copy()
        for _ in range(3):
            q = (np.abs(q) - np.float32(0.5)).astype(np.float32)
            if q[0] < q[1]:
                q[0], q[1] = q[1], q[0]
            q[2] = np.float32(q[2] * np.float32(1.5))
        acc = np.float32(0)
        for i in range(4):
            acc = np.float32(acc + np.float32(i) * np.float32(0.25))
        q = q[::-1]
        acc = np.float32(acc - np.float32(1.0))
        a = np.float32(np.sqrt(np.float32(np.float32(q[0] * q[0]) + np.float32(q[1] * q[1]))))
        b = max(q[0], q[2])
        return np.float32(np.float32((b if p[1] > 0 else a) + acc) + np.float32(2))

    got = run_wgsl(src, pts)
    want = np.array([ref(p) for p in pts], np.float32)
    assert f32_equal(got, want).all()


def test_glsl_features(built, tmp_path):
    glsl = textwrap.dedent("""\
        #version 450 core
        #define PI 3.14159265
        #define SQ(x) ((x)*(x))
        uniform float iTime;
        const float R = 0.75;
        float helper(in vec3 p, out float extra) { extra = p.y * 2.0; return length(p) - R; }
        float sdf(vec3 p) {
            float e;
            float d = helper(p, e);
            int n = 0;
            do { n++; d += 0.125; } while (n < 2);
            vec2 w = vec2(d, e);
            w.yx = w.xy;
            float s = (p.z > 0.0) ? SQ(w.x) : mod(w.y, 0.5);
            return s + sin(PI * iTime) + float(n) + max(vec3(p.x, 1, 2), 0.5).x;
        }
        void main() {}
        """)
    f = tmp_path / "t.frag"
    f.write_text(glsl)
    sh = s2m.Sdf3DShader.from_glsl_fragment_shader(f, "sdf")
    pts = points(3.0, 2000)

    def ref(p):
        p = p.astype(np.float32)
        e = np.float32(p[1] * np.float32(2))
        d = np.float32(np.sqrt(np.float32(np.float32(np.float32(p[0] * p[0]) + np.float32(p[1] * p[1])) + np.float32(p[2] * p[2]))) - np.float32(0.75))
        d = np.float32(np.float32(d + np.float32(0.125)) + np.float32(0.125))
        w = np.array([e, d], np.float32)  # w.yx = w.xy  ->  w = (old y, old x) = (e, d)
        if p[2] > 0:
            s = np.float32(w[0] * w[0])
        else:
            y = w[1]
            s = np.float32(y - np.float32(np.float32(0.5) * np.floor(np.float32(y / np.float32(0.5)))))
        out = np.float32(np.float32(np.float32(s + np.float32(0.0)) + np.float32(2)) + max(p[0], np.float32(0.5)))
        return out

    got = host_eval.eval_points(sh.lower_to_cuda(), pts)
    want = np.array([ref(p) for p in pts], np.float32)
    assert f32_equal(got, want).all()


def test_continue_to_break_rewrite(built, monkeypatch):
    """IR optimisation (frontend/optimize.cpp): an idempotent `P; if (c) continue;` loop prefix that does
    not depend on the loop counter turns `continue` into `break`; anything else is left alone, and the
    values never change"""
    ok = "fn sdf3d(p: vec3f) -> f32 { var r = 0.0; var z = p; for (var i = 0; i < 6; i++) { r = length(z); if (r > 2.0) { continue; } z = z * 1.7 + p; } return r; }"
    acc = "fn sdf3d(p: vec3f) -> f32 { var r = 0.0; var z = p; for (var i = 0; i < 6; i++) { r = r + length(z); if (r > 2.0) { continue; } z = z * 1.7 + p; } return r; }"
    cnt = "fn sdf3d(p: vec3f) -> f32 { var r = 0.0; var z = p; for (var i = 0; i < 6; i++) { r = length(z) + f32(i); if (r > 2.0) { continue; } z = z * 1.7 + p; } return r; }"
    outer = "fn sdf3d(p: vec3f) -> f32 { var r = 0.0; var z = p; var i = 0; for (; i < 6; i++) { r = length(z); if (r > 2.0) { continue; } z = z * 1.7 + p; } return r + f32(i); }"
    pts = points(4.0, 5000)
    for src, rewritten in ((ok, True), (acc, False), (cnt, False), (outer, False)):
        monkeypatch.delenv("S2M_NO_IR_OPT", raising=False)
        opt = s2m.Sdf3DShader.from_source(src).lower_to_cuda()
        monkeypatch.setenv("S2M_NO_IR_OPT", "1")
        plain = s2m.Sdf3DShader.from_source(src).lower_to_cuda()
        assert "continue;" in plain and "break;" not in plain
        assert ("break;" in opt) == rewritten
        assert f32_equal(host_eval.eval_points(opt, pts), host_eval.eval_points(plain, pts)).all()
    monkeypatch.delenv("S2M_NO_IR_OPT", raising=False)
    assert "break;" in load_example_shader("mandelbulb").lower_to_cuda()


def test_guarded_loop_rotation(built, monkeypatch):
    """IR optimisation (frontend/optimize.cpp: rotate_guarded_loop): `for (..) { P; if (c) break; R }` with a prefix of plain
    assignments becomes `P; if (!c) for (;;) { R; step; if (!cond) break; P; if (c) break; }` -- what the compiler hoists
    out of the loop for R's sake then sits behind the first test.  Loops whose rest `continue`s, whose prefix declares a
    variable, or that have no prefix are left alone; the values never change (NaN and all)"""
    rot = "fn sdf3d(p: vec3f) -> f32 { var r = 0.0; var z = p; var d = 1.0; for (var i = 0; i < 6; i++) { r = length(z); if (r > 2.0) { break; } d = d * r * 2.0 + 1.0; z = z * 1.7 + p; } return 0.5 * log(r) * r / d; }"
    cont = "fn sdf3d(p: vec3f) -> f32 { var r = 0.0; var z = p; for (var i = 0; i < 6; i++) { r = length(z); if (r > 2.0) { break; } if (z.x > 1.0) { z = z * 0.5; continue; } z = z * 1.7 + p; } return r; }"
    decl = "fn sdf3d(p: vec3f) -> f32 { var r = 0.0; var z = p; for (var i = 0; i < 6; i++) { let q = length(z); if (q > 2.0) { break; } r = r + q; z = z * 1.7 + p; } return r; }"
    bare = "fn sdf3d(p: vec3f) -> f32 { var r = 0.0; var z = p; for (var i = 0; i < 6; i++) { if (r > 2.0) { break; } r = r + length(z); z = z * 1.7 + p; } return r; }"
    nested = "fn sdf3d(p: vec3f) -> f32 { var r = 0.0; var z = p; for (var i = 0; i < 4; i++) { r = length(z); if (r > 3.0) { break; } for (var j = 0; j < 3; j++) { if (z.y > 0.5) { continue; } z.y = z.y + 0.3; } z = z * 1.3 + p; } return r; }"
    zero = "fn sdf3d(p: vec3f) -> f32 { var r = 7.0; var z = p; for (var i = 0; i < i32(p.x); i++) { r = length(z); if (r > 2.0) { break; } z = z * 1.7 + p; } return r; }"
    pts = points(4.0, 5000)
    for src, rotated in ((rot, True), (cont, False), (decl, False), (bare, False), (nested, True), (zero, True)):
        monkeypatch.delenv("S2M_NO_LOOP_ROTATION", raising=False)
        opt = s2m.Sdf3DShader.from_source(src).lower_to_cuda()
        monkeypatch.setenv("S2M_NO_LOOP_ROTATION", "1")
        plain = s2m.Sdf3DShader.from_source(src).lower_to_cuda()
        assert "for (; ; )" not in plain
        assert ("for (; ; )" in opt) == rotated, src
        assert f32_equal(host_eval.eval_points(opt, pts), host_eval.eval_points(plain, pts)).all(), src
    monkeypatch.delenv("S2M_NO_LOOP_ROTATION", raising=False)
    sh = load_example_shader("mandelbulb")
    assert "for (; ; )" in sh.lower_to_cuda() and "for (; ; )" in sh.lower_to_cuda_packed()


def test_sin_cos_pairing(built, monkeypatch):
    """IR optimisation (frontend/optimize.cpp: pair_sin_cos): sin(e) / cos(e) of the same pure argument
    become one sincos_pair(e) unless a variable of e is written in between; values never change"""
    cases = [
        # (source, number of pairs expected)
        ("fn sdf3d(p: vec3f) -> f32 { let a = p.x * 40000.0; let s = sin(a); let c = cos(a); return s * p.y + c * p.z; }", 1),
        ("fn sdf3d(p: vec3f) -> f32 { return sin(p.x * 3.0) * p.y + cos(p.x * 3.0) * p.z + sin(p.y) ; }", 1),
        ("fn sdf3d(p: vec3f) -> f32 { var a = p.x; let s = sin(a); a = a + 1.0; let c = cos(a); return s + c; }", 0),
        ("fn sdf3d(p: vec3f) -> f32 { var a = p.x; let s = sin(a); if (p.y > 0.0) { a = 2.0 * a; } let c = cos(a); return s + c; }", 0),
        ("fn sdf3d(p: vec3f) -> f32 { var q = p; q.x = sin(q.y) + cos(q.y); q.y = cos(q.z) - sin(q.z) + cos(q.x) * sin(q.x); return length(q) - 1.0; }", 3),
        ("fn rot(v: vec2f, a: f32) -> vec2f { return vec2f(cos(a) * v.x - sin(a) * v.y, sin(a) * v.x + cos(a) * v.y); } fn sdf3d(p: vec3f) -> f32 { let q = rot(p.xy, 1.0e6 * p.z); return length(vec3f(q, p.z)) - 1.0; }", 1),
        ("fn sdf3d(p: vec3f) -> f32 { var acc = 0.0; for (var i = 0; i < 3; i++) { let a = p.x * f32(i + 1); acc = acc + sin(a) * cos(a); } return acc + sin(p.y) + cos(p.z); }", 1),
    ]
    pts = points(4.0, 4000)
    for src, n_pairs in cases:
        monkeypatch.delenv("S2M_NO_IR_OPT", raising=False)
        opt = s2m.Sdf3DShader.from_source(src).lower_to_cuda()
        monkeypatch.setenv("S2M_NO_IR_OPT", "1")
        plain = s2m.Sdf3DShader.from_source(src).lower_to_cuda()
        assert "f_sincos_pair" not in plain
        assert opt.count("= f_sincos_pair(") == n_pairs, src
        assert f32_equal(host_eval.eval_points(opt, pts), host_eval.eval_points(plain, pts)).all(), src
    monkeypatch.delenv("S2M_NO_IR_OPT", raising=False)
    assert load_example_shader("mandelbulb").lower_to_cuda().count("= f_sincos_pair(") == 2


def test_mutable_globals_switch_structs_arrays(built, tmp_path):
    """f2: module-scope variables that functions assign, switch, structs, fixed-size arrays and dynamic
    indexing -- the same program in WGSL and in GLSL (through the GLSL -> WGSL -> CUDA route), checked
    against a numpy transcription"""
    wgsl = textwrap.dedent("""\
        const N = 3;
        struct Hit { d: f32, id: i32, }
        struct Scene { spheres: array<vec4f, N>, best: Hit, }
        const RADII = array<f32, N>(0.5, 0.25, 0.75);
        var<private> evals: i32;
        var<private> bias: f32 = 0.125;
        fn closer(a: Hit, b: Hit) -> Hit { evals = evals + 1; if (a.d < b.d) { return a; } return b; }
        fn make_scene() -> Scene {
          var s: Scene;
          for (var i = 0; i < N; i++) { s.spheres[i] = vec4f(f32(i) - 1.0, 0.0, 0.0, RADII[i]); }
          s.best = Hit(1e9, -1);
          return s;
        }
        fn weight(id: i32) -> f32 {
          var w = 0.0;
          switch id {
            case 0: { w = 1.0; }
            case 1, 2: { w = 2.0; if (id == 2) { break; } w = w + 0.5; }
            default: { w = -1.0; }
          }
          return w;
        }
        fn sdf3d(p: vec3f) -> f32 {
          var sc = make_scene();
          for (var i = 0; i < N; i++) {
            let sp = sc.spheres[i];
            sc.best = closer(sc.best, Hit(length(p - sp.xyz) - sp.w, i));
          }
          bias = bias * f32(evals);
          var v = p;
          v[sc.best.id] = v[sc.best.id] * 2.0;
          let m = mat3x3f(v, p, v + p);
          let far = array(p.x, p.y, p.z)[sc.best.id + 7];    // out of range: nearest element
          return sc.best.d + 0.001 * weight(sc.best.id) + 0.01 * m[sc.best.id].y + 0.0001 * v[2] + 0.00001 * far + bias;
        }
        """)
    glsl = textwrap.dedent("""\
        #version 450 core
        const int N = 3;
        struct Hit { float d; int id; };
        struct Scene { vec4 spheres[N]; Hit best; };
        const float RADII[N] = float[N](0.5, 0.25, 0.75);
        int evals;
        float bias = 0.125;
        Hit closer(Hit a, Hit b) { evals++; if (a.d < b.d) return a; return b; }
        void make_scene(out Scene s) {
          for (int i = 0; i < N; i++) s.spheres[i] = vec4(float(i) - 1.0, 0.0, 0.0, RADII[i]);
          s.best = Hit(1e9, -1);
        }
        float weight(int id) {
          float w = 0.0;
          switch (id) {
            case 0: w = 1.0; break;
            case 1:
            case 2: w = 2.0; if (id == 2) break; w += 0.5; break;
            default: w = -1.0;
          }
          return w;
        }
        float sdf(vec3 p) {
          Scene sc;
          make_scene(sc);
          for (int i = 0; i < N; i++) {
            vec4 sp = sc.spheres[i];
            sc.best = closer(sc.best, Hit(length(p - sp.xyz) - sp.w, i));
          }
          bias *= float(evals);
          vec3 v = p;
          v[sc.best.id] *= 2.0;
          mat3 m = mat3(v, p, v + p);
          float far = float[](p.x, p.y, p.z)[sc.best.id + 7];
          return sc.best.d + 0.001 * weight(sc.best.id) + 0.01 * m[sc.best.id].y + 0.0001 * v[2] + 0.00001 * far + bias;
        }
        void main() {}
        """)
    f = np.float32

    def expect(p):
        best_d, best_id = f(1e9), -1
        for i, rad in enumerate((f(0.5), f(0.25), f(0.75))):
            q = (p - np.array([f(i) - f(1), 0, 0], np.float32)).astype(np.float32)
            d = f(np.sqrt(f(f(q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]))) - rad
            if not (best_d < d):
                best_d, best_id = f(d), i
        bias = f(f(0.125) * f(3))
        v = p.copy()
        v[best_id] = v[best_id] * f(2)
        m = [v, p, (v + p).astype(np.float32)]
        w = (f(1.0), f(2.5), f(2.0))[best_id]
        r = f(best_d + f(f(0.001) * w))
        r = f(r + f(f(0.01) * m[best_id][1]))
        r = f(r + f(f(0.0001) * v[2]))
        r = f(r + f(f(0.00001) * p[2]))
        return f(r + bias)

    pts = points(4.0, 1500)
    want = np.array([expect(p) for p in pts], np.float32)
    frag = tmp_path / "agg.frag"
    frag.write_text(glsl)
    shaders = {"wgsl": s2m.Sdf3DShader.from_source(wgsl), "glsl": s2m.Sdf3DShader.from_glsl_fragment_shader(frag, "sdf")}
    for lang, sh in shaders.items():
        cuda = sh.lower_to_cuda()
        assert "struct S_Hit" in cuda and "struct S2mState" in cuda and "switch (" in cuda and "s2m_array<vec4, 3>" in cuda
        assert f32_equal(host_eval.eval_points(cuda, pts), want).all(), lang
        assert sh.create_shader_module(None).cubin_size > 0   # NVRTC accepts it for sm_100a
    assert "struct Scene {" in shaders["glsl"].source and "var<private> evals: i32;" in shaders["glsl"].source
    # GLSL fall-through between non-empty cases: the following statements are parsed into the falling case (test_glsl_switch_fall_through)
    frag.write_text("#version 450\nfloat sdf(vec3 p) { float w = 0.0; switch (int(p.x)) { case 0: w = 1.0; case 1: w += 2.0; break; } return w; }\nvoid main() {}\n")
    fall = s2m.Sdf3DShader.from_glsl_fragment_shader(frag, "sdf")
    assert host_eval.eval_points(fall.lower_to_cuda(), np.array([[0.5, 0, 0], [1.5, 0, 0], [2.5, 0, 0]], np.float32)).tolist() == [3.0, 2.0, 0.0]


def test_integer_vectors_and_bit_functions(built, tmp_path):
    """f2: uvec / ivec arithmetic with wrap-around, shifts, bit operations, integer min/max/clamp/abs/sign,
    WGSL-defined division by zero, floatBitsToUint / bitcast, GLSL relational functions, mix with a bool
    selector, refract -- an integer-hash value-noise SDF in GLSL and a bit-twiddling one in WGSL, each
    against a numpy transcription"""
    glsl = textwrap.dedent("""\
        #version 450 core
        uvec3 pcg3d(uvec3 v) {
          v = v * 1664525u + 1013904223u;
          v.x += v.y * v.z; v.y += v.z * v.x; v.z += v.x * v.y;
          v ^= v >> 16u;
          v.x += v.y * v.z; v.y += v.z * v.x; v.z += v.x * v.y;
          return v;
        }
        vec3 hash33(vec3 p) {
          uvec3 q = pcg3d(uvec3(ivec3(floor(p)) + 1000));
          return vec3(q) * (1.0 / float(0xffffffffu));
        }
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
          float e = float(m - 127) * 0.001;
          float bump = 0.05 * vnoise(p * 4.0) + (any(neg) ? 0.0 : 0.01) + (all(not(neg)) ? 0.002 : 0.0);
          ivec3 c = clamp(ivec3(p * 2.0), ivec3(-2), ivec3(2));
          return length(q) - 1.0 + bump + e + 0.001 * float(abs(c.x) + max(c.y, c.z) + sign(c.z)) + 0.0001 * float(7 / (c.x - c.x));
        }
        void main() {}
        """)
    f, u32 = np.float32, np.uint32

    def pcg3d(v):
        v = (v * u32(1664525) + u32(1013904223)).astype(u32)
        v[0] += v[1] * v[2]; v[1] += v[2] * v[0]; v[2] += v[0] * v[1]
        v ^= v >> u32(16)
        v[0] += v[1] * v[2]; v[1] += v[2] * v[0]; v[2] += v[0] * v[1]
        return v

    def hash33x(p):
        q = pcg3d((np.floor(p).astype(np.int32) + np.int32(1000)).astype(u32))
        return f(f(q[0]) * f(f(1.0) / f(4294967295.0)))

    def vnoise(p):
        i, fr = np.floor(p).astype(np.float32), (p - np.floor(p)).astype(np.float32)
        u = ((fr * fr).astype(np.float32) * (f(3.0) - (f(2.0) * fr).astype(np.float32)).astype(np.float32)).astype(np.float32)
        acc = f(0.0)
        for k in range(8):
            o = np.array([k & 1, (k >> 1) & 1, (k >> 2) & 1])
            w = np.where(o == 1, u, (f(1.0) - u).astype(np.float32)).astype(np.float32)
            acc = f(acc + f(f(f(hash33x((i + o.astype(np.float32)).astype(np.float32)) * w[0]) * w[1]) * w[2]))
        return acc

    def expect(p):
        neg = p < 0
        q = np.where(neg, -p, p).astype(np.float32)
        m = int((q[:1].view(u32)[0] >> u32(23)) & u32(0xff))
        e = f(f(m - 127) * f(0.001))
        bump = f(f(f(0.05) * vnoise((p * f(4.0)).astype(np.float32))) + (f(0.0) if neg.any() else f(0.01)))
        bump = f(bump + (f(0.002) if (~neg).all() else f(0.0)))
        c = np.clip(np.trunc((p * f(2.0)).astype(np.float32)).astype(np.int32), -2, 2)
        r = f(f(np.sqrt(f(f(q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]))) - f(1.0))
        r = f(f(r + bump) + e)
        r = f(r + f(f(0.001) * f(abs(int(c[0])) + max(int(c[1]), int(c[2])) + int(np.sign(c[2])))))
        return f(r + f(f(0.0001) * f(7)))       # 7 / 0 = 7 (WGSL rule)

    with np.errstate(over="ignore"):
        pts = points(4.0, 600)
        want = np.array([expect(p) for p in pts], np.float32)
    frag = tmp_path / "hash.frag"
    frag.write_text(glsl)
    sh = s2m.Sdf3DShader.from_glsl_fragment_shader(frag, "sdf")
    assert "vec3<u32>" in sh.source and "bitcast<u32>(" in sh.source
    cuda = sh.lower_to_cuda()
    assert "uvec3 u_pcg3d(uvec3" in cuda
    assert f32_equal(host_eval.eval_points(cuda, pts), want).all()
    assert sh.create_shader_module(None).cubin_size > 0
    wgsl = textwrap.dedent("""\
        fn mantissa_bits(x: f32) -> u32 { return bitcast<u32>(x) & 0x7fffffu; }
        fn sdf3d(p: vec3f) -> f32 {
          let b = bitcast<vec3<u32>>(abs(p));
          let e = vec3<i32>((b >> vec3<u32>(23u)) & vec3<u32>(255u)) - vec3<i32>(127);
          let lo = mantissa_bits(p.x) >> 20u;
          let s = select(vec3<i32>(1), vec3<i32>(-1), p < vec3f(0.0));
          let t = (e * s) % vec3<i32>(3, 0, 2);
          let one = bitcast<f32>(0x3f800000u);
          return length(p) - one + 0.01 * f32(t.x + t.y + t.z) + 0.001 * f32(lo) + 0.0001 * f32(max(e.x, min(e.y, e.z)) << 2u)
                 + refract(normalize(p + 0.1), vec3f(0.0, 1.0, 0.0), 0.9).x * 1e-3 + faceForward(p, vec3f(0.0, 0.0, 1.0), p).z * 1e-3;
        }
        """)

    def expect2(p):
        b = np.abs(p).astype(np.float32).view(u32)
        e = ((b >> u32(23)) & u32(255)).astype(np.int32) - np.int32(127)
        lo = int((p[:1].view(u32)[0] & u32(0x7fffff)) >> u32(20))
        s = np.where(p < 0, -1, 1)
        es = e * s
        t = [int(np.fmod(es[0], 3)), 0, int(np.fmod(es[2], 2))]    # truncated remainder; x % 0 = 0
        ln = f(np.sqrt(f(f(p[0] * p[0] + p[1] * p[1]) + p[2] * p[2])))
        r = f(ln - f(1.0))
        r = f(r + f(f(0.01) * f(t[0] + t[1] + t[2])))
        r = f(r + f(f(0.001) * f(lo)))
        r = f(r + f(f(0.0001) * f(int(max(e[0], min(e[1], e[2]))) << 2)))
        i = (p + f(0.1)).astype(np.float32)
        i = (i / f(np.sqrt(f(f(i[0] * i[0] + i[1] * i[1]) + i[2] * i[2])))).astype(np.float32)
        eta = f(0.9)
        d = f(f(f(0.0) * i[0] + f(1.0) * i[1]) + f(0.0) * i[2])
        k = f(f(1.0) - f(f(eta * eta) * f(f(1.0) - f(d * d))))
        rx = f(0.0) if k < 0 else f(f(eta * i[0]) - f(f(f(eta * d) + f(np.sqrt(k))) * f(0.0)))
        r = f(r + f(rx * f(1e-3)))
        ff = p[2] if f(f(f(p[0] * f(0.0) + p[1] * f(0.0)) + p[2] * f(1.0))) < 0 else -p[2]
        return f(r + f(f(ff) * f(1e-3)))

    pts2 = points(4.0, 600)
    want2 = np.array([expect2(p) for p in pts2], np.float32)
    cuda2 = s2m.Sdf3DShader.from_source(wgsl).lower_to_cuda()
    assert f32_equal(host_eval.eval_points(cuda2, pts2), want2).all()


def test_matrix_inverse_and_logical_xor(built, tmp_path):
    """GLSL inverse() for mat2/mat3/mat4 (adjugate / determinant, pinned operation order), determinant(mat4), ^^"""
    glsl = textwrap.dedent("""\
        #version 450 core
        float sdf(vec3 p) {
          mat2 m = mat2(2.0, 1.0, 0.0, 4.0);
          mat3 r = mat3(1.0, 0.5, 0.0,  0.0, 2.0, 0.25,  0.1, 0.0, 1.5);
          mat4 t = mat4(1.0, 0.0, 0.0, 0.0,  0.0, 2.0, 0.0, 0.0,  0.0, 0.5, 1.0, 0.0,  0.3, 0.2, 0.1, 1.0);
          bool a = p.x > 0.0, b = p.y > 0.0;
          vec4 h = inverse(t) * vec4(p, 1.0);
          return (inverse(m) * p.xy).x + (inverse(r) * p).z + h.x + h.w + ((a ^^ b) ? 1.0 : 0.0) + determinant(t);
        }
        void main() {}
        """)
    frag = tmp_path / "inv.frag"
    frag.write_text(glsl)
    sh = s2m.Sdf3DShader.from_glsl_fragment_shader(frag, "sdf")
    pts = points(4.0, 500)
    got = host_eval.eval_points(sh.lower_to_cuda(), pts)
    m = np.array([[2, 0], [1, 4]], np.float64)
    r = np.array([[1, 0, 0.1], [0.5, 2, 0], [0, 0.25, 1.5]])
    t = np.array([[1, 0, 0, 0.3], [0, 2, 0.5, 0.2], [0, 0, 1, 0.1], [0, 0, 0, 1]])
    want = []
    for p in pts.astype(np.float64):
        h = np.linalg.inv(t) @ np.append(p, 1)
        want.append((np.linalg.inv(m) @ p[:2])[0] + (np.linalg.inv(r) @ p)[2] + h[0] + h[3] + float((p[0] > 0) != (p[1] > 0)) + np.linalg.det(t))
    assert np.abs(got - np.array(want)).max() < 1e-5
    assert sh.create_shader_module(None).cubin_size > 0


def test_shadertoy_style_raymarcher(built):
    """tests/data/shadertoy_raymarcher.glsl -- structs with ?:, mat2 rotation, smooth min, mod repetition,
    value noise, swizzle stores, #define constants, mainImage + ShaderToy uniforms -- through
    from_shadertoy_source; values against a float64 numpy transcription (tolerance, not bits)"""
    code = open(os.path.join(ROOT, "tests", "data", "shadertoy_raymarcher.glsl")).read()
    sh = s2m.Sdf3DShader.from_shadertoy_source(code, "map")
    assert "fn mapHit(" in sh.source and "struct Hit {" in sh.source
    pts = points(4.0, 1500)
    got = host_eval.eval_points(sh.lower_to_cuda(), pts)

    f = np.float32
    F = lambda *v: np.array(v, np.float32)

    def fract(x): return (x - np.floor(x)).astype(np.float32)
    def mix(a, b, t): return f(f(a * f(f(1) - t)) + f(b * t))
    def length(v):
        acc = f(v[0] * v[0])
        for c in v[1:]:
            acc = f(acc + f(c * c))
        return f(np.sqrt(acc))
    def dot(a, b):
        acc = f(a[0] * b[0])
        for x, y in zip(a[1:], b[1:]):
            acc = f(acc + f(x * y))
        return acc
    def hash_(p):
        p = fract((p * f(0.3183099) + f(0.1)).astype(np.float32))
        p = (p * f(17.0)).astype(np.float32)
        return fract(f(f(f(p[0] * p[1]) * p[2]) * f(f(p[0] + p[1]) + p[2])))
    def noise(x):
        i, fr = np.floor(x).astype(np.float32), fract(x)
        fr = ((fr * fr).astype(np.float32) * (f(3) - (f(2) * fr).astype(np.float32)).astype(np.float32)).astype(np.float32)
        h = lambda a, b, c: hash_((i + F(a, b, c)).astype(np.float32))
        return mix(mix(mix(h(0, 0, 0), h(1, 0, 0), fr[0]), mix(h(0, 1, 0), h(1, 1, 0), fr[0]), fr[1]),
                   mix(mix(h(0, 0, 1), h(1, 0, 1), fr[0]), mix(h(0, 1, 1), h(1, 1, 1), fr[0]), fr[1]), fr[2])
    def box(p, s):
        p = (np.abs(p) - s).astype(np.float32)
        return f(length(np.maximum(p, f(0))) + min(max(p[0], max(p[1], p[2])), f(0)))
    def smin(a, b, k):
        h = min(max(f(f(0.5) + f(f(f(0.5) * f(b - a)) / k)), f(0)), f(1))
        return f(mix(b, a, h) - f(f(k * h) * f(f(1) - h)))
    s_, c_ = f(0.479425550), f(0.877582550)       # sin / cos of fl(0 * 0.2 + 0.5) (iTime = 0), rounded to f32

    def expect(p):
        q = p.copy()
        x, z = q[0], q[2]                         # q.xz *= mat2(c, -s, s, c): row vector times matrix
        q[0], q[2] = f(f(x * c_) + f(z * -s_)), f(f(x * s_) + f(z * c_))
        d = box(q, F(0.5, 0.5, 0.5))
        rp = p.copy()
        t0 = f(rp[0] + f(1))
        rp[0] = f(f(t0 - f(f(2) * np.floor(f(t0 / f(2))))) - f(1))
        t = (rp - F(0, 0.8, 0)).astype(np.float32)
        tor = f(length(F(f(length(F(t[0], t[2])) - f(0.4)), t[1])) - f(0.1))
        d = d if d < tor else tor
        a_, b_ = F(-1, 0, 0), F(1, 0.5, 0.3)
        ab, ap = (b_ - a_).astype(np.float32), (p - a_).astype(np.float32)
        tt = min(max(f(dot(ab, ap) / dot(ab, ab)), f(0)), f(1))
        cap = f(length((p - (a_ + (tt * ab).astype(np.float32)).astype(np.float32)).astype(np.float32)) - f(0.15))
        d = d if d < cap else cap
        d = smin(d, f(length((p - F(0, -0.6, 0)).astype(np.float32)) - f(0.4)), f(0.2))
        return f(d + f(f(0.03) * noise((p * f(6.0)).astype(np.float32))))

    want = np.array([expect(p) for p in pts], np.float32)
    assert np.abs(got - want).max() < 2e-6   # the only non-pinned step above is rounding sin/cos(0.5) by hand
    assert sh.create_shader_module(None).cubin_size > 0


def test_shadertoy_idioms_swizzled_inout_argument(built):
    """tests/data/shadertoy_idioms.glsl: pModPolar(q.xz, 6.) -- an inout parameter fed with a swizzle --
    plus function-like macros, a global written per call, a constant array, float loop counters,
    while(true)/break, folding with swizzle swaps; against a float64 numpy transcription"""
    code = open(os.path.join(ROOT, "tests", "data", "shadertoy_idioms.glsl")).read()
    sh = s2m.Sdf3DShader.from_shadertoy_source(code, "map")
    assert "pModPolar(&swz_1, 6f);" in sh.source and "q.z = swz_1.y;" in sh.source   # standard WGSL, no &q.xz
    pts = points(4.0, 1500)
    got = host_eval.eval_points(sh.lower_to_cuda(), pts)

    def rbox(p, b, r):
        q = np.abs(p) - b
        return np.linalg.norm(np.maximum(q, 0)) + min(max(q[0], max(q[1], q[2])), 0) - r
    def hexprism(p, h):
        k = np.array([-0.8660254, 0.5, 0.57735])
        p = np.abs(p)
        p[:2] = p[:2] - 2.0 * min(k[:2] @ p[:2], 0.0) * k[:2]
        d0 = np.linalg.norm(p[:2] - np.array([np.clip(p[0], -k[2] * h[0], k[2] * h[0]), h[0]])) * np.sign(p[1] - h[0])
        d = np.array([d0, p[2] - h[1]])
        return min(max(d[0], d[1]), 0.0) + np.linalg.norm(np.maximum(d, 0.0))
    def mod(x, y): return x - y * np.floor(x / y)
    def fractal(p):
        s_ = 1.0
        for i in range(5):
            p = np.abs(p) - np.array([0.6, 0.4, 0.3]) * s_
            if p[0] < p[1]: p[[0, 1]] = p[[1, 0]]
            if p[0] < p[2]: p[[0, 2]] = p[[2, 0]]
            a = 0.3 + i * 0.1
            y, z = p[1], p[2]                      # p.yz *= mat2(c, s, -s, c): row vector times matrix
            p[1], p[2] = y * np.cos(a) + z * np.sin(a), y * -np.sin(a) + z * np.cos(a)
            s_ *= 0.6
        return rbox(p, np.full(3, 0.1 * s_ * 4.0), 0.01)
    def expect(p):
        d = 1e10
        q = p.copy()
        ang = 6.2831853 / 6.0
        a = np.arctan2(q[2], q[0]) + ang / 2
        r = np.hypot(q[0], q[2])
        a = mod(a, ang) - ang / 2
        q[0], q[2] = np.cos(a) * r, np.sin(a) * r
        d = min(d, hexprism(q - np.array([1.2, 0, 0]), (0.2, 0.3)))
        t = np.float32(0.0)
        while t < 1.0:
            d = min(d, np.linalg.norm(p - np.array([0, float(t) - 0.5, 0])) - 0.15 * (1 - float(t)))
            t = np.float32(t + np.float32(0.34))
        for k in range(3):
            d = min(d, np.linalg.norm(p - np.eye(3)[k] * 0.9) - 0.1)
        d = min(d, fractal(p * 1.3) / 1.3)
        rr = mod(p + 0.75, 1.5) - 0.75
        return max(d, -(np.linalg.norm(rr) - 0.2))     # gTime = 0

    want = np.array([expect(p) for p in pts.astype(np.float64)])
    err = np.abs(got - want)
    assert np.quantile(err, 0.99) < 2e-5 and err.max() < 1e-3
    assert sh.create_shader_module(None).cubin_size > 0


def test_glsl_comma_in_for_header(built, tmp_path):
    """for (int i = 0, j = 5; i < j; i++, j--): several declarators and a comma-separated continuing part
    become a WGSL loop with a continuing block, so `continue` still advances both counters"""
    frag = tmp_path / "comma.frag"
    frag.write_text(textwrap.dedent("""\
        #version 450 core
        float sdf(vec3 p) {
          float r = 0.0;
          for (int i = 0, j = 5; i < j; i++, j--) { if (i == 1) continue; r += p.x * float(i) + p.y * float(j); }
          int k; float w;
          for (k = 0, w = 1.0; k < 3; ++k, w *= 0.5) r += w * p.z;
          return r;
        }
        void main() {}
        """))
    sh = s2m.Sdf3DShader.from_glsl_fragment_shader(frag, "sdf")
    assert "continuing {" in sh.source
    pts = points(4.0, 300)
    want = []
    for p in pts.astype(np.float64):
        r, i, j = 0.0, 0, 5
        while i < j:
            if i != 1:
                r += p[0] * i + p[1] * j
            i, j = i + 1, j - 1
        r += (1.0 + 0.5 + 0.25) * p[2]
        want.append(r)
    assert np.abs(host_eval.eval_points(sh.lower_to_cuda(), pts) - np.array(want)).max() < 1e-5


def test_glsl_initializer_lists(built, tmp_path):
    """GLSL 4.20 brace initializers for arrays (sized and unsized), structs (nested) and vectors"""
    frag = tmp_path / "init.frag"
    frag.write_text(textwrap.dedent("""\
        #version 450 core
        struct S { float r; vec3 c; float w[2]; };
        const float K[] = {0.5, 0.25, 0.125};
        float sdf(vec3 p) {
          float a[3] = {1.0, 2.0, 3.0};
          S s = {0.75, {0.1, 0.2, 0.3}, {2.0, 4.0}};
          vec2 v = {p.x, p.y};
          return length(p - s.c) - s.r + a[1] * K[2] + s.w[1] * 0.01 + v.y * 0.001;
        }
        void main() {}
        """))
    sh = s2m.Sdf3DShader.from_glsl_fragment_shader(frag, "sdf")
    pts = points(4.0, 300)
    want = [np.linalg.norm(p - np.array([0.1, 0.2, 0.3])) - 0.75 + 2 * 0.125 + 0.04 + p[1] * 0.001 for p in pts.astype(np.float64)]
    assert np.abs(host_eval.eval_points(sh.lower_to_cuda(), pts) - np.array(want)).max() < 1e-5


def test_glsl_side_effects_inside_expressions(built, tmp_path):
    """`d += e = map(p)`, `w[i++]`, `while (i++ < n)`, `for (; k++ < 3;)`, `do ... while (++j < 3)`, `a = b = c`,
    `if ((e = f(p)) > x)`: each effect becomes its own statement in front of the statement that contains it
    (loop conditions: in front of every test); effects on the conditional side of && / || / ?: are refused"""
    frag = tmp_path / "fx.frag"
    frag.write_text(textwrap.dedent("""\
        #version 450 core
        float map(vec3 p) { return length(p) - 1.0; }
        float sdf(vec3 p) {
          float d = 0.0, e, acc = 0.0;
          int i = 0, n = 0;
          float w[4] = float[4](1.0, 2.0, 4.0, 8.0);
          d += e = map(p);
          acc += w[i++] + w[i++];
          while (i++ < 4) acc += 0.5;
          for (int k = 0; k++ < 3;) { if (k == 2) continue; n += k; }
          int j = 0;
          do { acc += 0.25; } while (++j < 3);
          float a, b2;
          a = b2 = p.x * 2.0;
          if ((e = abs(p.y)) > 0.5) acc += e;
          return d + acc + float(n) + a + b2 + float(i) * 0.01 + float(j) * 0.001;
        }
        void main() {}
        """))
    sh = s2m.Sdf3DShader.from_glsl_fragment_shader(frag, "sdf")
    assert "var _post1: i32 = i;" in sh.source and "break if !((j < 3i));" in sh.source
    pts = points(4.0, 300)
    want = []
    for p in pts.astype(np.float64):
        acc = 1 + 2 + 2 * 0.5 + 3 * 0.25        # w[0] + w[1]; i = 2 -> two passes of the while; three of the do
        e = abs(p[1])
        if e > 0.5:
            acc += e
        want.append((np.linalg.norm(p) - 1.0) + acc + (1 + 3) + 4 * p[0] + 5 * 0.01 + 3 * 0.001)
    assert np.abs(host_eval.eval_points(sh.lower_to_cuda(), pts) - np.array(want)).max() < 1e-5
    for bad in ("float sdf(vec3 p) { float e = 0.0; if (p.x > 0.0 && (e = p.y) > 0.0) return e; return 1.0; }",
                "float sdf(vec3 p) { int i = 0; if (p.y > 9.0) return 0.0; else if ((p.x > 0.0 ? float(i++) : 1.0) > 0.5) return 2.0; return float(i); }",
                "int g = 0; float h = float(g++); float sdf(vec3 p) { return h; }"):
        frag.write_text("#version 450\n" + bad + "\nvoid main() {}\n")
        with pytest.raises(s2m.S2mError) as e:
            s2m.Sdf3DShader.from_glsl_fragment_shader(frag, "sdf")
        assert "UNSUPPORTED" in str(e.value)
    # an effect inside an arm of ?: where a statement can be issued: the arm becomes a branch of an if / else
    frag.write_text("#version 450\nfloat sdf(vec3 p) { int i = 3; float r = p.x > 0.0 ? float(i++) : 1.0; return r + 10.0 * float(i); }\nvoid main() {}\n")
    sh = s2m.Sdf3DShader.from_glsl_fragment_shader(frag, "sdf")
    assert host_eval.eval_points(sh.lower_to_cuda(), np.array([[1, 0, 0], [-1, 0, 0]], np.float32)).tolist() == [43.0, 31.0]


def test_modf_and_ldexp(built, tmp_path):
    """GLSL modf(x, out whole) / WGSL modf(x).fract|.whole (whole = trunc(x)), ldexp(x, e) = x * 2^e"""
    frag = tmp_path / "modf.frag"
    frag.write_text("#version 450\nfloat sdf(vec3 p) { float i; float f = modf(p.x * 3.0, i); vec3 w; vec3 g = modf(p * 2.0, w); "
                    "return f + 10.0 * i + g.y + 100.0 * w.z + ldexp(p.y, 3) + ldexp(1.0, int(p.z)); }\nvoid main() {}\n")
    pts = points(4.0, 300)
    P = pts.astype(np.float64)
    got = host_eval.eval_points(s2m.Sdf3DShader.from_glsl_fragment_shader(frag, "sdf").lower_to_cuda(), pts)
    want = (P[:, 0] * 3 - np.trunc(P[:, 0] * 3)) + 10 * np.trunc(P[:, 0] * 3) + (P[:, 1] * 2 - np.trunc(P[:, 1] * 2)) + 100 * np.trunc(P[:, 2] * 2) \
        + P[:, 1] * 8 + 2.0 ** np.trunc(P[:, 2])
    assert np.abs(got - want).max() < 1e-4
    wgsl = "fn sdf3d(p: vec3f) -> f32 { return modf(p.x * 3.0).fract + 10.0 * modf(p.x * 3.0).whole + modf(p * 2.0).fract.y + ldexp(p.y, 3); }"
    got = host_eval.eval_points(s2m.Sdf3DShader.from_source(wgsl).lower_to_cuda(), pts)
    want = (P[:, 0] * 3 - np.trunc(P[:, 0] * 3)) + 10 * np.trunc(P[:, 0] * 3) + (P[:, 1] * 2 - np.trunc(P[:, 1] * 2)) + P[:, 1] * 8
    assert np.abs(got - want).max() < 1e-4


def test_matrices(built, tmp_path):
    """mat2/mat3 (GLSL) and mat2x2f/mat3x3<f32> (WGSL): constructors, m*v, v*m, m*m, m[i], transpose"""
    glsl = textwrap.dedent("""\
        #version 450 core
        mat2 rot(float a) { float c = cos(a), s = sin(a); return mat2(c, -s, s, c); }
        float sdf(vec3 p) {
            p.xz *= rot(0.7);
            p.xy = rot(-0.3) * p.xy;
            mat3 m = mat3(vec3(1.0, 0.5, 0.0), vec3(0.0, 1.0, 0.25), vec3(0.1, 0.0, 1.0));
            vec3 q = transpose(m) * p + m[1] * 0.5;
            mat3 mm = m * mat3(2.0);
            return length(q * mm) - determinant(m);
        }
        void main() {}
        """)
    f = tmp_path / "m.frag"
    f.write_text(glsl)
    sh = s2m.Sdf3DShader.from_glsl_fragment_shader(f, "sdf")
    assert "mat2x2<f32>" in sh.source and "mat3x3<f32>" in sh.source
    pts = points(3.0, 1000)
    F = np.float32

    def rot(a):
        c, s = F(np.cos(F(a))), F(np.sin(F(a)))
        return np.array([[c, -s], [s, c]], F)  # columns: (c,-s), (s,c)

    def mv(cols, v):  # sum_j cols[j]*v[j], left to right, f32
        acc = cols[0] * v[0]
        for j in range(1, len(v)):
            acc = (acc + cols[j] * v[j]).astype(F)
        return acc.astype(F)

    def vm(v, cols):
        out = []
        for c in cols:
            acc = F(v[0] * c[0])
            for j in range(1, len(v)):
                acc = F(acc + F(v[j] * c[j]))
            out.append(acc)
        return np.array(out, F)

    def ref(p):
        p = p.astype(F).copy()
        r1 = rot(0.7)
        p[[0, 2]] = vm(p[[0, 2]], r1)           # p.xz = p.xz * rot
        p[[0, 1]] = mv(rot(-0.3), p[[0, 1]])    # p.xy = rot * p.xy
        m = np.array([[1.0, 0.5, 0.0], [0.0, 1.0, 0.25], [0.1, 0.0, 1.0]], F)  # m[j] = column j
        mt = m.T.copy()                          # transpose: columns of mt = rows of m
        q = (mv(mt, p) + m[1] * F(0.5)).astype(F)
        two = np.diag(np.full(3, 2.0, F)).astype(F)
        mm = np.array([mv(m, two[j]) for j in range(3)], F)
        v = vm(q, mm)
        ln = F(np.sqrt(F(F(F(v[0] * v[0]) + F(v[1] * v[1])) + F(v[2] * v[2]))))
        cr = np.array([F(F(m[1][1] * m[2][2]) - F(m[1][2] * m[2][1])), F(F(m[1][2] * m[2][0]) - F(m[1][0] * m[2][2])),
                       F(F(m[1][0] * m[2][1]) - F(m[1][1] * m[2][0]))], F)
        det = F(F(F(m[0][0] * cr[0]) + F(m[0][1] * cr[1])) + F(m[0][2] * cr[2]))
        return F(ln - det)

    got = host_eval.eval_points(sh.lower_to_cuda(), pts)
    want = np.array([ref(p) for p in pts], F)
    # cos/sin here are the engine's pinned functions; numpy's may differ in the last ulp
    assert np.allclose(got, want, rtol=2e-6, atol=2e-6)
    wgsl = "fn sdf3d(p: vec3f) -> f32 { let m = mat2x2f(vec2f(0.0, 1.0), vec2f(-1.0, 0.0)); let n = mat2x2<f32>(1.0, 2.0, 3.0, 4.0); let q = (m * n) * p.xy; return q.x + n[1].y + (p.xy * m).y; }"
    v = run_wgsl(wgsl, [[1.0, 2.0, 0.0]])[0]
    # m*n columns: m*(1,2) = (-2,1), m*(3,4) = (-4,3);  (m*n)*(1,2) = (-2-8, 1+6) = (-10, 7);  (p.xy*m).y = dot((1,2),(-1,0)) = -1
    assert v == np.float32(-10.0 + 4.0 - 1.0)


SHADERTOY_CODE = """
// a typical ShaderToy image pass: helpers, overloads, uniforms, mainImage
#define PI 3.14159265
float hash(float n) { return fract(sin(n) * 43758.5453); }
float hash(vec2 p) { return hash(p.x + p.y * 57.0); }
mat2 rot(float a) { float c = cos(a), s = sin(a); return mat2(c, -s, s, c); }
float map(vec3 p) {
    p.xz *= rot(0.4 + iTime);
    vec3 q = abs(p) - vec3(0.6, 0.4, 0.5);
    float box = length(max(q, 0.0)) + min(max(q.x, max(q.y, q.z)), 0.0);
    return min(box, length(p - vec3(0.0, 0.7, 0.0)) - 0.35) + 0.0 * hash(p.xy) + 0.0 * hash(1.0);
}
void mainImage(out vec4 fragColor, in vec2 fragCoord) {
    vec2 uv = fragCoord / iResolution.xy;
    fragColor = vec4(uv, 0.5 + 0.5 * sin(iTime), 1.0);
}
"""


def test_shadertoy_source_path(built):
    """from_shadertoy_api minus the network: uniform block, mainImage removal, --shadertoy-sdf name"""
    sh = s2m.Sdf3DShader.from_shadertoy_source(SHADERTOY_CODE, "map")
    src = sh.source
    assert "fn mainImage(" not in src and "fn main(" not in src and "@fragment" not in src
    assert "fn hash(" in src and "fn hash_1(" in src, "GLSL overloads get distinct WGSL names"
    assert "fn sdf3d(p: vec3<f32>) -> f32 { return map(p); }" in src and "fn sdf3d_normal(" in src
    pts = points(3.0, 2000)
    got = host_eval.eval_points(sh.lower_to_cuda(), pts)

    def ref(p):
        F = np.float32
        a = F(0.4)
        c, s_ = F(np.cos(a)), F(np.sin(a))
        x = F(F(p[0] * c) + F(p[2] * F(-s_)))   # p.xz * mat2(c,-s,s,c): (dot(p.xz,(c,-s)), dot(p.xz,(s,c)))
        z = F(F(p[0] * s_) + F(p[2] * c))
        q = np.array([abs(x) - F(0.6), abs(p[1]) - F(0.4), abs(z) - F(0.5)], F)
        m = np.maximum(q, F(0))
        box = F(F(np.sqrt(F(F(F(m[0] * m[0]) + F(m[1] * m[1])) + F(m[2] * m[2])))) + min(max(q[0], max(q[1], q[2])), F(0)))
        d = np.array([x, p[1] - F(0.7), z], F)
        sph = F(F(np.sqrt(F(F(F(d[0] * d[0]) + F(d[1] * d[1])) + F(d[2] * d[2])))) - F(0.35))
        return min(box, sph)

    want = np.array([ref(p) for p in pts], np.float32)
    assert np.allclose(got, want, rtol=3e-6, atol=3e-6)
    with pytest.raises(s2m.S2mError) as e:
        s2m.Sdf3DShader.from_shadertoy_source(SHADERTOY_CODE, "sdf")
    assert e.value.kind == "MISSING_SDF"
    with pytest.raises(s2m.S2mError) as e:
        s2m.Sdf3DShader.from_shadertoy_source("float map(vec3 p) { return p.x; }", "map")  # no mainImage
    assert e.value.kind == "SHADER" and "mainImage" in str(e.value)


def test_glsl_preprocessor_conditionals(built, tmp_path):
    glsl = textwrap.dedent("""\
        #version 450 core
        #define QUALITY 2
        #define USE_SPHERE
        #if QUALITY > 1 && defined(USE_SPHERE)
        #define RADIUS 0.75
        #elif QUALITY == 1
        #define RADIUS 0.5
        #else
        #define RADIUS 0.25
        #endif
        #ifdef NOT_DEFINED
        this is not GLSL and must be skipped
        #endif
        #ifndef NOT_DEFINED
        #define OFFSET 0.125   // trailing comment
        #endif
        #undef QUALITY
        #if defined QUALITY
        #define OFFSET 100.0
        #endif
        float sdf(vec3 p) { return length(p) - RADIUS + OFFSET; }
        void main() {}
        """)
    f = tmp_path / "pp.frag"
    f.write_text(glsl)
    sh = s2m.Sdf3DShader.from_glsl_fragment_shader(f, "sdf")
    v = host_eval.eval_points(sh.lower_to_cuda(), np.array([[2.0, 0.0, 0.0]], np.float32))[0]
    assert v == np.float32(np.float32(2.0) - np.float32(0.75)) + np.float32(0.125)
    f.write_text("#version 450 core\n#if 1\nfloat sdf(vec3 p) { return p.x; }\nvoid main() {}\n")
    with pytest.raises(s2m.S2mError) as e:
        s2m.Sdf3DShader.from_glsl_fragment_shader(f, "sdf")
    assert e.value.kind == "PARSE" and "unterminated #if" in str(e.value)


def test_packed_lane_uniformity_analysis(built, tmp_path):
    """The packed emitter keeps values that cannot differ between the two lanes as plain floats.  The
    cases that decide a variable's kind: assignment of a varying value, by-reference parameters (one object
    behind two names), module-scope state, aggregates (always pairs), float -> int conversions and bit casts
    (lanes must agree).  The analysis is monovariant: a function has ONE signature, so a parameter is a
    pair as soon as one call site passes a varying value (helpers used only with constants stay scalar)."""
    src = textwrap.dedent("""\
        struct Acc { sum: f32, n: i32, }
        var<private> seen: f32;
        fn scale(v: f32, k: f32) -> f32 { return v * k + 0.125; }            // v: varying at its call site, k: uniform
        fn scale_u(v: f32, k: f32) -> f32 { return v * k + 0.125; }          // only ever called with uniform values
        fn bump(a: ptr<function, f32>, by: f32) { *a = *a + by; seen = seen + by; }
        fn bump_u(a: ptr<function, f32>, by: f32) { *a = *a + by; seen = seen + by; }
        fn sdf3d(p: vec3f) -> f32 {
            var gain = 2.0;                       // uniform all the way
            gain = scale_u(gain, 1.5);
            var u = 0.5;                          // starts uniform, becomes varying through the pointer
            bump(&u, p.x);
            var w = 0.25;                         // stays uniform: its function only sees uniform values
            bump_u(&w, gain);
            var acc: Acc;
            acc.sum = scale(p.y, gain);
            acc.n = i32(floor(p.z * 3.0));        // varying float -> int
            var t = vec3f(gain, w, 1.0);          // uniform vector ...
            t.y = t.y + p.z;                      // ... until a component becomes varying
            let bits = bitcast<u32>(p.x) >> 31u;  // sign bit: lanes must agree
            var r = length(p + t) - u + acc.sum * 0.01 + f32(acc.n) * 0.001 + seen * 0.5 + f32(bits);
            if (w > 3.0) { r = r + 1.0; }         // comparison of uniform values
            return r;
        }""")
    sh = s2m.Sdf3DShader.from_source(src)
    packed = sh.lower_to_cuda_packed()
    assert "float u_gain" in packed and "float u_w" in packed          # never varying
    assert "pf u_u" in packed and "pvec3 u_t" in packed                # became varying
    assert "pf u_scale(S2mState& G, pf u_v, float u_k)" in packed
    assert "float u_scale_u(S2mState& G, float u_v, float u_k)" in packed
    assert "void u_bump(S2mState& G, pf& u_a, pf u_by)" in packed and "void u_bump_u(S2mState& G, float& u_a, float u_by)" in packed
    assert "pf u_seen;" in packed                                        # module-scope state written with a varying value
    assert "p_f2int(G," in packed and "p_bits_u(G," in packed
    assert "((u_w_" in packed and ") > (3.0f))" in packed               # plain comparison, no lane check
    pts = points(4.0, 20_000)

    def ref(p):
        f = np.float32
        gain = f(f(f(2.0) * f(1.5)) + f(0.125))
        u = f(f(0.5) + p[0]); seen = p[0]
        w = f(f(0.25) + gain); seen = f(seen + gain)
        acc_sum = f(f(p[1] * gain) + f(0.125))
        n = int(np.floor(f(p[2] * f(3.0))))
        t = np.array([gain, f(w + p[2]), f(1.0)], f)
        q = (p + t).astype(f)
        ln = f(np.sqrt(f(f(f(q[0] * q[0]) + f(q[1] * q[1])) + f(q[2] * q[2]))))
        bits = 1.0 if np.signbit(p[0]) else 0.0
        r = f(f(f(f(f(ln - u) + f(acc_sum * f(0.01))) + f(f(n) * f(0.001))) + f(seen * f(0.5))) + f(bits))
        return f(r + f(1.0)) if w > 3.0 else r

    got = host_eval.eval_points(sh.lower_to_cuda(), pts)   # (also runs the packed form on the same points)
    want = np.array([ref(p) for p in pts], np.float32)
    assert f32_equal(got, want).all()
    a, b = pts, np.roll(pts, 1, axis=0)
    lo, hi, dv = host_eval.eval_pairs(packed, a, b)
    assert f32_equal(lo, want).all()
    assert (f32_equal(hi, np.roll(want, 1)) | dv).all()
    # the lanes disagree exactly when floor(3z) differs or the bit patterns of x differ (the bit cast itself
    # is where a float turns into an integer; that only its sign bit is used afterwards is not seen)
    expect_dv = (np.floor(a[:, 2] * np.float32(3)) != np.floor(b[:, 2] * np.float32(3))) | (a[:, 0].view(np.uint32) != b[:, 0].view(np.uint32))
    assert np.array_equal(dv, expect_dv)


def test_packed_form_was_cross_checked(built):
    """tests/conftest.py lowers every shader of this file to its packed f32x2 form as well, and
    host_eval.eval_points compares the two lane by lane (lane lo always; lane hi unless the lanes
    disagreed).  This test only makes sure that this happened for the bulk of the suite."""
    st = host_eval.PACKED_STATS
    print(f"packed cross-check: {len(st['shaders'])} shaders, {st['pairs']} pairs, {st['disagreed']} with disagreeing lanes, "
          f"{len(st['unpackable'])} shaders without a packed form (matrices)")
    assert len(st["shaders"]) >= 20
    assert st["pairs"] > 100_000
    assert len(st["unpackable"]) <= 10


def test_comparison_of_two_abstract_constants(built):
    """`(-2.5) == (-1.0)` has a concrete (bool) type but abstract operands: it folds in abstract precision
    (found by tests/test_frontend_fuzz.py: the literals used to reach the emitter unconverted)"""
    sh = s2m.Sdf3DShader.from_source("fn sdf3d(p: vec3f) -> f32 { return select(p.y, p.z, ((-2.5) == (-1.0))) + select(1.0, p.x, 1 < 2) "
                                     "+ select(0.0, 8.0, 0.1 + 0.2 == 0.3); }")
    pts = points(2.0, 500)
    got = host_eval.eval_points(sh.lower_to_cuda(), pts)
    # 0.1 + 0.2 == 0.3 is false in abstract-float (f64) arithmetic, true in f32
    assert f32_equal(got, (pts[:, 1] + pts[:, 0]).astype(np.float32)).all()


def test_glsl_names_that_wgsl_reserves(built, tmp_path):
    """GLSL identifiers that are WGSL keywords, reserved words or predeclared types get a `_` suffix in the
    WGSL text (naga's namer does the same), so the dumped WGSL is valid and parses back"""
    glsl = textwrap.dedent("""\
        #version 450 core
        struct type { float fn; float ok; };
        float f16(vec3 target) { return target.x * 2.0; }
        float sdf(vec3 p) {
            float filter = p.y, let = 0.5, f32 = p.z, f32_ = 1.0;
            type self = type(filter, let);
            return f16(p) + self.fn * self.ok + f32 * f32_;
        }
        void main() {}
        """)
    wgsl = s2m.convert_glsl_to_wgsl(glsl)
    for name in ("struct type_ ", "fn_: f32", "fn f16_(target_: vec3<f32>)", "var filter_:", "var let_:", "var f32__:", "var self_: type_", "self_.fn_ * self_.ok"):
        assert name in wgsl, (name, wgsl)
    f = tmp_path / "r.frag"
    f.write_text(glsl)
    sh = s2m.Sdf3DShader.from_glsl_fragment_shader(f, "sdf")
    pts = points(2.0, 500)
    got = host_eval.eval_points(sh.lower_to_cuda(), pts)
    x, y, z = (pts[:, i] for i in range(3))
    want = ((x * np.float32(2) + (y * np.float32(0.5)).astype(np.float32)).astype(np.float32) + (z * np.float32(1)).astype(np.float32)).astype(np.float32)
    assert f32_equal(got, want).all()


def test_shading_code_with_textures_and_derivatives_is_left_out(built):
    """a ShaderToy image pass is shading code around the distance function: functions that need textures,
    channel inputs or screen-space derivatives (and their callers) are left out of the module with a note;
    it is an error only when the SDF itself is among them"""
    code = textwrap.dedent("""\
        float map(vec3 p) { return length(p) - 0.75; }
        vec3 shade(vec3 p, vec2 uv) { vec3 t = texture(iChannel0, uv).xyz; return t * fwidth(p.x); }
        float aa(float d) { return d / fwidth(d); }
        vec3 render(vec3 ro, vec3 rd, vec2 uv) { float t = 0.0; for (int i = 0; i < 64; i++) { t += map(ro + rd * t); } return shade(ro + rd * t, uv); }
        float after(vec3 p) { return map(p) * 2.0; }
        void mainImage(out vec4 fragColor, in vec2 fragCoord) { vec2 uv = fragCoord / iResolution.xy; fragColor = vec4(render(vec3(0, 0, -3), vec3(uv, 1), uv), 1.0); }
        """)
    sh = s2m.Sdf3DShader.from_shadertoy_source(code, "map")
    src = sh.source
    assert "fn map(" in src and "fn after(" in src and "fn shade(" not in src and "fn render(" not in src and "fn aa(" not in src
    assert "// left out: fn shade -- " in src and "iChannel0" in src and "// left out: fn render -- calls 'shade'" in src and "// left out: fn aa -- " in src
    pts = points(2.0, 300)
    want = (np.sqrt(((pts[:, 0] * pts[:, 0]).astype(np.float32) + (pts[:, 1] * pts[:, 1]).astype(np.float32)).astype(np.float32)
                    + (pts[:, 2] * pts[:, 2]).astype(np.float32)).astype(np.float32) - np.float32(0.75)).astype(np.float32)
    assert f32_equal(host_eval.eval_points(sh.lower_to_cuda(), pts), want).all()
    assert sh.create_shader_module(None).cubin_size > 0
    for name, why in (("aa", "fwidth"), ("render", "calls 'shade'"), ("shade", "iChannel0")):
        with pytest.raises(s2m.S2mError) as e:
            s2m.Sdf3DShader.from_shadertoy_source(code, name)
        assert e.value.kind == "UNSUPPORTED" and why in str(e.value), str(e.value)
    with pytest.raises(s2m.S2mError) as e:
        s2m.Sdf3DShader.from_shadertoy_source(code, "nope")
    assert e.value.kind == "MISSING_SDF"
    # a syntax error is still an error, wherever it is
    with pytest.raises(s2m.S2mError) as e:
        s2m.Sdf3DShader.from_shadertoy_source(code + "\nfloat broken(vec3 p) { return p.x +; }\n", "map")
    assert e.value.kind == "PARSE"


def test_glsl_isnan_isinf(built):
    src = ("#version 450 core\nfloat sdf(vec3 p) { float big = p.x * 3.0e38; float n = (big - big) * p.y; vec2 v = vec2(big * 4.0, p.z);\n"
           "  return float(isnan(n)) + 2.0 * float(isinf(big * 4.0)) + 4.0 * float(isinf(p.y)) + 8.0 * float(any(isinf(v))) + 16.0 * float(any(isnan(vec3(n, 1.0, 2.0)))); }\nvoid main() {}\n")
    sh = s2m.Sdf3DShader.from_source(src, s2m.SRC_GLSL_FRAGMENT, "sdf")
    pts = np.array([[1, 1, 1], [2, 0.5, 0], [0.1, 1, 1], [-3, 2, 1], [0, 0, 0]], np.float32)
    assert host_eval.eval_points(sh.lower_to_cuda(), pts).tolist() == [10.0, 27.0, 0.0, 27.0, 0.0]
    assert sh.create_shader_module(None).cubin_size > 0


def test_glsl_idioms_from_shadertoy_code(built):
    """braced switch cases that end in return / break, do-while over a local array, a float loop counter,
    matrix column and element stores, a dynamically indexed vector store, integer xor / and, while with continue"""
    src = textwrap.dedent("""\
        #version 450 core
        #define N 4
        float shape(vec3 p, int kind) {
            switch (kind) {
                case 0: return p.x - 1.0;
                case 1: { vec3 d = abs(p) - vec3(0.75); return max(max(d.x, d.y), d.z); }
                case 2:
                case 3: { if (p.y > 0.0) { return p.y * 0.5; } else { break; } }
                default: break;
            }
            return 8.0;
        }
        float sdf(vec3 p) {
            float arr[N];
            for (int i = 0; i < N; i++) arr[i] = shape(p, i);
            float d = 100.0;
            int k = 0;
            do { d = min(d, arr[k]); k++; } while (k < N);
            for (float f = 0.; f < 1.; f += .25) d = min(d, abs(p.z - f) + 0.5);
            mat3 m = mat3(1.0);
            m[1] = vec3(0., 2., 0.);
            m[2][0] = 0.5;
            vec3 w = m * p;
            w[k % 3] += 0.25;
            ivec2 ij = ivec2(floor(p.xy * 4.0));
            d += float((ij.x ^ ij.y) & 1) * 0.125;
            int n = 0;
            while (n < 3) { if (d > float(n) * 0.25) { n += 2; continue; } n++; }
            return d + w.x + w.y * 0.5 + w.z * 0.25 + float(n);
        }
        void main() {}
        """)
    sh = s2m.Sdf3DShader.from_source(src, s2m.SRC_GLSL_FRAGMENT, "sdf")
    F = np.float32

    def shape(p, kind):
        if kind == 0:
            return F(p[0] - F(1))
        if kind == 1:
            d = np.abs(p) - F(0.75)
            return max(max(d[0], d[1]), d[2])
        if kind in (2, 3) and p[1] > 0:
            return F(p[1] * F(0.5))
        return F(8)

    def ref(p):
        arr = [shape(p, i) for i in range(4)]
        d = F(100)
        for k in range(4):
            d = min(d, arr[k])
        f = F(0)
        while f < 1:
            d = min(d, F(F(abs(F(p[2] - f))) + F(0.5)))
            f = F(f + F(0.25))
        # m = columns (1,0,0), (0,2,0), (0.5,0,1);  w = m * p = c0*p.x + c1*p.y + c2*p.z, summed left to right
        w = [F(F(F(F(1) * p[0]) + F(F(0) * p[1])) + F(F(0.5) * p[2])),
             F(F(F(F(0) * p[0]) + F(F(2) * p[1])) + F(F(0) * p[2])),
             F(F(F(F(0) * p[0]) + F(F(0) * p[1])) + F(F(1) * p[2]))]
        w[4 % 3] = F(w[4 % 3] + F(0.25))
        ij = np.floor(p[:2] * F(4)).astype(np.int32)
        d = F(d + F(F((int(ij[0]) ^ int(ij[1])) & 1) * F(0.125)))
        n = 0
        while n < 3:
            if d > F(F(n) * F(0.25)):
                n += 2
                continue
            n += 1
        return F(F(F(F(d + w[0]) + F(w[1] * F(0.5))) + F(w[2] * F(0.25))) + F(n))

    pts = points(3.0, 1500)
    got = host_eval.eval_points(sh.lower_to_cuda(), pts)
    want = np.array([ref(p) for p in pts], np.float32)
    assert f32_equal(got, want).all()
    assert sh.create_shader_module(None).cubin_size > 0


def test_glsl_whole_operand_equality_comma_statements_matrix_resize_uniform_blocks(built):
    """GLSL semantics WGSL does not share: == / != on vectors and matrices give ONE bool; the comma operator at
    statement level; matN(matM) keeps the upper-left block and fills with the identity; sampler uniforms are
    accepted (their users are left out); uniform blocks with and without an instance name read as zero"""
    src = textwrap.dedent("""\
        #version 450 core
        layout(binding = 1) uniform sampler2D tex0, tex1[2];
        layout(std140, binding = 2) uniform Params { float t; vec3 c; float w[2]; };
        layout(std140) uniform Scene { vec4 sphere; int n; } scene;
        vec3 albedo(vec2 uv) { return texture(tex0, uv).rgb; }
        float sdf(vec3 p) {
            float a, b; int i = 0;
            a = 1., b = 2.;
            a += p.x, b *= p.y, i++;
            mat2 m = mat2(1.), n = mat2(1., 0., 0., p.z);
            float r = (p.xy == vec2(0.5, 0.25)) ? 100. : 0.;
            r += (p.xy != p.yx) ? 10. : 0.;
            r += (m == n) ? 1000. : 0.;
            r += (m != n) ? 5. : 0.;
            mat4 big = mat4(vec4(1., 2., 3., 4.), vec4(5., 6., 7., 8.), vec4(9., 10., 11., 12.), vec4(13., 14., 15., 16.));
            mat3 n3 = mat3(big);
            mat4 back = mat4(mat2(n3));
            r += (n3 * vec3(0., 1., 0.)).z * 1e4 + (back * vec4(0., 0., 1., 1.)).z * 1e5 + (back * vec4(1., 0., 0., 0.)).y * 1e6;
            return r + a + b + float(i) + t + c.x + w[1] + scene.sphere.w + float(scene.n);
        }
        void main() {}
        """)
    sh = s2m.Sdf3DShader.from_source(src, s2m.SRC_GLSL_FRAGMENT, "sdf")
    assert "// left out: fn albedo -- " in sh.source and "sampler2D uniform 'tex0'" in sh.source
    assert "var<private> scene: Scene;" in sh.source and "var<private> c: vec3<f32>;" in sh.source
    pts = np.array([[0.5, 0.25, 1.0], [0.5, 0.5, 2.0], [1, 2, 1], [0.5, 0.25, 3]], np.float32)
    got = host_eval.eval_points(sh.lower_to_cuda(), pts)
    # (n3 * e_y).z = 7, (back * (0,0,1,1)).z = 1, (back * e_x).y = 2
    assert got.tolist() == [2170000.0 + v for v in (1113.0, 8.5, 1017.0, 118.0)]
    assert sh.create_shader_module(None).cubin_size > 0


def test_byte_order_mark_and_crlf_line_ends(built, tmp_path):
    """files saved on other platforms: UTF-8 BOM, \\r\\n line ends (also after a preprocessor line continuation), tabs, non-ASCII comments"""
    f = tmp_path / "w.sdf3d"
    f.write_bytes(b"\xef\xbb\xbfuse sdf3d::*;\r\n// caf\xc3\xa9 \xe2\x80\x94 comment\r\nfn sdf3d(p: vec3f) -> f32 {\r\n\treturn sdf3d_torus(p, vec2f(0.75, 0.25));\r\n}\r\n")
    a = s2m.Sdf3DShader.from_path(f)
    g = tmp_path / "g.frag"
    g.write_bytes(b"\xef\xbb\xbf#version 450 core\r\n#define R 0.25 // caf\xc3\xa9\r\n#define LONG(a) \\\r\n   ((a) + 0.5)\r\nfloat sdf(vec3 p) {\r\n"
                  b"\tvec2 q = vec2(length(p.xz) - (LONG(0.0) + 0.25), p.y);\r\n\treturn length(q) - R;\r\n}\r\nvoid main() {}\r\n")
    b = s2m.Sdf3DShader.from_glsl_fragment_shader(g, "sdf")
    pts = points(2.0, 2000)
    va, vb = host_eval.eval_points(a.lower_to_cuda(), pts), host_eval.eval_points(b.lower_to_cuda(), pts)
    assert f32_equal(va, vb).all() and np.isfinite(va).all()


def test_from_shadertoy_api_and_json(built):
    """Sdf3DShader::from_shadertoy_api (shader.rs:110-144) with the HTTP client replaced: URL and key as the
    reference builds them (shadertoy.rs:126-131), the code of all render passes concatenated (:126-132 of impl
    Shader), `Error` responses -> ShaderError, a failed request -> RequestError, a missing SDF -> MissingSdf"""
    import json
    common = "float sphere(vec3 p, float r) { return length(p) - r; }\n"
    image = "float sdf(vec3 p) { return sphere(p, 0.75); }\nvoid mainImage(out vec4 c, in vec2 u) { c = vec4(sdf(vec3(u, iTime))); }\n"
    response = {"Shader": {"ver": "0.1", "info": {"id": "DldfR7", "name": "ball", "username": "someone"},
                           "renderpass": [{"inputs": [], "outputs": [], "code": common, "name": "Common", "type": "common"},
                                          {"inputs": [], "outputs": [], "code": image, "name": "Image", "type": "image"}]}}
    seen = []

    def fetch(url):
        seen.append(url)
        return json.dumps(response).encode()

    sh = s2m.Sdf3DShader.from_shadertoy_api("DldfR7", "sdf", fetch=fetch)
    assert seen == ["https://www.shadertoy.com/api/v1/shaders/DldfR7?key=rdnjhn"]
    assert sh.info == {"name": "ball", "username": "someone"}
    assert "fn sphere(" in sh.source and "fn sdf3d(p: vec3<f32>) -> f32 { return sdf(p); }" in sh.source and "mainImage" not in sh.source
    pts = points(2.0, 500)
    want = (np.sqrt(((pts[:, 0] * pts[:, 0]).astype(np.float32) + (pts[:, 1] * pts[:, 1]).astype(np.float32)).astype(np.float32)
                    + (pts[:, 2] * pts[:, 2]).astype(np.float32)).astype(np.float32) - np.float32(0.75)).astype(np.float32)
    assert f32_equal(host_eval.eval_points(sh.lower_to_cuda(), pts), want).all()
    with pytest.raises(s2m.S2mError) as e:
        s2m.Sdf3DShader.from_shadertoy_api("DldfR7", "distance", fetch=fetch)
    assert e.value.kind == "MISSING_SDF"
    with pytest.raises(s2m.S2mError) as e:
        s2m.Sdf3DShader.from_shadertoy_api("nope", fetch=lambda url: b'{"Error": "Shader not found"}')
    assert e.value.kind == "SHADER" and "Shader not found" in str(e.value)

    def offline(url):
        raise OSError("network unreachable")

    with pytest.raises(s2m.S2mError) as e:
        s2m.Sdf3DShader.from_shadertoy_api("DldfR7", fetch=offline)
    assert e.value.kind == "REQUEST" and "network unreachable" in str(e.value)
    for bad in ("[1, 2]", "{", '{"Shader": {"renderpass": [{"code": "abc}]}}', '{"Other": 1}', ""):
        with pytest.raises(s2m.S2mError) as e:
            s2m.Sdf3DShader.from_shadertoy_json(bad)
        assert e.value.kind == "SHADER"
    # JSON escapes in the code string: \n \t \" \\ \/ \u00e9 and a surrogate pair, nested objects with their own "code" keys are not passes
    body = ('{"Shader": {"info": {"name": "esc \\u00e9", "username": "u", "tags": ["a", "b"]}, "ver": "0.1", "renderpass": ['
            '{"inputs": [{"id": 1, "sampler": {"filter": "linear", "code": "not code"}}], "outputs": [], '
            '"code": "// caf\\u00e9 \\ud83d\\ude00 \\"quoted\\" a\\/b\\nfloat sdf(vec3 p) {\\n\\treturn length(p) - 0.75; // back\\\\slash\\n}\\nvoid mainImage(out vec4 c, in vec2 u) { c = vec4(0.0); }\\n", "name": "Image", "type": "image"}]}}')
    sh = s2m.Sdf3DShader.from_shadertoy_json(body)
    assert sh.info == {"name": "esc \u00e9", "username": "u"}
    assert f32_equal(host_eval.eval_points(sh.lower_to_cuda(), pts), want).all()


def test_integer_bit_builtins_and_abstract_int_folding(built):
    """countOneBits / reverseBits / countLeadingZeros / countTrailingZeros / firstLeadingBit / firstTrailingBit (WGSL) =
    bitCount / bitfieldReverse / findMSB / findLSB (GLSL, int results); abs / sign / min / max / clamp of abstract-int
    constants stay abstract-int (tests/test_frontend_fuzz.py found them turning into abstract-float)"""
    w = s2m.Sdf3DShader.from_source(
        "fn sdf3d(p: vec3f) -> f32 { let u = bitcast<u32>(p.x); let i = bitcast<i32>(p.y);\n"
        "  let a = f32(countOneBits(u)) + 64.0 * f32(countLeadingZeros(u)) + 4096.0 * f32(countTrailingZeros(i));\n"
        "  let b = f32(firstLeadingBit(i)) + 64.0 * f32(firstTrailingBit(i)) + 4096.0 * f32(reverseBits(u) >> 24u);\n"
        "  let c = (i32(p.z) ^ sign((-2147483647 - 1))) + abs(-5) + min(1, 2) + clamp(7, 0, 3) + i32(firstLeadingBit(0u) == 0xffffffffu) + firstLeadingBit(vec2i(-1, 5)).x;\n"
        "  return a + 1.0e6 * b + 1.0e9 * f32(c); }")
    g = s2m.Sdf3DShader.from_source(
        "#version 450 core\nfloat sdf(vec3 p) { uint u = floatBitsToUint(p.x); int i = floatBitsToInt(p.y);\n"
        "  float a = float(bitCount(u)) + 64.0 * float(31 - findMSB(u)) + 4096.0 * float(findLSB(i) < 0 ? 32 : findLSB(i));\n"
        "  float b = float(findMSB(i)) + 64.0 * float(findLSB(i)) + 4096.0 * float(bitfieldReverse(u) >> 24u);\n"
        "  int c = (int(p.z) ^ sign(-2147483647 - 1)) + abs(-5) + min(1, 2) + clamp(7, 0, 3) + int(findMSB(0u) == -1) + findMSB(ivec2(-1, 5)).x;\n"
        "  return a + 1.0e6 * b + 1.0e9 * float(c); }\nvoid main() {}\n", s2m.SRC_GLSL_FRAGMENT, "sdf")
    pts = np.concatenate([points(4.0, 1500), np.array([[0, 0, 0], [-0.0, -0.0, 1], [1, -1, -3], [np.inf, -np.inf, 2]], np.float32)])
    va, vb = host_eval.eval_points(w.lower_to_cuda(), pts), host_eval.eval_points(g.lower_to_cuda(), pts)
    assert f32_equal(va, vb).all()

    def ref(p):
        u, i = int(p[:1].view(np.uint32)[0]), int(p[1:2].view(np.int32)[0])
        iu = i & 0xffffffff
        clz = 32 - u.bit_length()
        ctz_i = 32 if iu == 0 else (iu & -iu).bit_length() - 1
        t = ~i if i < 0 else i
        flb = -1 if t == 0 else t.bit_length() - 1
        ftb = -1 if iu == 0 else ctz_i
        rev = int(format(u, "032b")[::-1], 2) >> 24
        F = np.float32
        a = F(F(F(bin(u).count("1")) + F(F(64) * F(clz))) + F(F(4096) * F(ctz_i)))
        b = F(F(F(flb) + F(F(64) * F(ftb))) + F(F(4096) * F(rev)))
        z = int(np.clip(np.trunc(np.float64(p[2])), -2**31, 2**31 - 1)) if np.isfinite(p[2]) else 0
        c = ((z ^ -1) + 5 + 1 + 3 + 1 + -1)
        c = (c + 2**31) % 2**32 - 2**31
        return F(F(a + F(F(1.0e6) * b)) + F(F(1.0e9) * F(c)))

    want = np.array([ref(p) for p in pts], np.float32)
    assert f32_equal(va, want).all()
    assert w.create_shader_module(None).cubin_size > 0 and g.create_shader_module(None).cubin_size > 0


def test_glsl_switch_fall_through(built):
    """a case that runs into the next label executes the following cases' statements up to the first break / return:
    the front-end parses those statements into the falling case again (the IR, like WGSL, has no fall-through)"""
    src = textwrap.dedent("""\
        #version 450 core
        float sdf(vec3 p) {
            int k = int(floor(p.x));
            float r = 0.;
            switch (k) {
                case 0: r = 1.; break;
                case 1: r = 2.;            // falls into default
                default: r += 3.;          // falls into case 5
                case 5: { float t = 10.; r += t; }
                case 6: case 7: r += 100.; if (p.y > 0.) break; r += 1000.;
                case 8: r += 1e4; break;
                case 9: r = (p.y > 0. ? 7. : 8.);
            }
            return r;
        }
        void main() {}
        """)
    sh = s2m.Sdf3DShader.from_source(src, s2m.SRC_GLSL_FRAGMENT, "sdf")
    xs = np.arange(-1, 11)
    pts = np.stack([xs + 0.5, np.where(xs % 2 == 0, 1.0, -1.0), np.zeros(len(xs))], 1).astype(np.float32)

    def ref(k, y):
        r = 0.0
        for step in range({0: 0, 1: 1, 5: 3, 6: 4, 7: 4, 8: 5, 9: 6}.get(k, 2), 7):
            if step == 0:
                return 1.0
            if step == 1:
                r = 2.0
            if step == 2:
                r += 3.0
            if step == 3:
                r += 10.0
            if step == 4:
                r += 100.0
                if y > 0:
                    return r
                r += 1000.0
            if step == 5:
                return r + 1e4
            if step == 6:
                r = 7.0 if y > 0 else 8.0
        return r

    got = host_eval.eval_points(sh.lower_to_cuda(), pts)
    assert got.tolist() == [ref(int(np.floor(p[0])), p[1]) for p in pts]
    assert sh.create_shader_module(None).cubin_size > 0


def test_extract_and_insert_bits(built):
    """extractBits / insertBits (GLSL bitfieldExtract / bitfieldInsert) with WGSL's clamping of offset and count, signed
    extraction sign-extends; offset and count stay scalars in the WGSL text when the value is a vector"""
    w = ("fn sdf3d(p: vec3f) -> f32 { let u = bitcast<u32>(p.x); let i = bitcast<i32>(p.x); let o = u32(abs(p.y) * 40.0); let c = u32(abs(p.z) * 40.0);\n"
         " return f32(extractBits(u, o, c) & 0xffffu) + 65536.0 * f32(extractBits(i, o, c) & 0xff) + f32(insertBits(u, 0x5a5a5a5au, o, c) >> 20u) * 0.0001"
         " + f32(extractBits(vec2u(u, 7u), 1u, 2u).y) * 1e8 + f32(insertBits(i, -1, o, c) & 0xfff) * 1e-8; }")
    g = ("#version 450 core\nfloat sdf(vec3 p) { uint u = floatBitsToUint(p.x); int i = floatBitsToInt(p.x); int o = int(abs(p.y) * 40.0); int c = int(abs(p.z) * 40.0);\n"
         " return float(bitfieldExtract(u, o, c) & 0xffffu) + 65536.0 * float(bitfieldExtract(i, o, c) & 0xff) + float(bitfieldInsert(u, 0x5a5a5a5au, o, c) >> 20u) * 0.0001"
         " + float(bitfieldExtract(uvec2(u, 7u), 1, 2).y) * 1e8 + float(bitfieldInsert(i, -1, o, c) & 0xfff) * 1e-8; }\nvoid main() {}\n")
    pts = np.random.default_rng(3).uniform(-1, 1, (3000, 3)).astype(np.float32)
    a, b = s2m.Sdf3DShader.from_source(w), s2m.Sdf3DShader.from_source(g, s2m.SRC_GLSL_FRAGMENT, "sdf")
    assert "extractBits(vec2<u32>(u, 7u), 1u, 2u)" in b.source
    va, vb = host_eval.eval_points(a.lower_to_cuda(), pts), host_eval.eval_points(b.lower_to_cuda(), pts)
    assert f32_equal(va, vb).all()

    def clamp_oc(o, c):
        o = min(o, 32)
        return o, min(c, 32 - o)

    def ext_u(e, o, c):
        o, c = clamp_oc(o, c)
        return 0 if c == 0 else (e >> o) & ((1 << c) - 1)

    def ext_i(e, o, c):
        o, c = clamp_oc(o, c)
        if c == 0:
            return 0
        v = ((e & 0xffffffff) >> o) & ((1 << c) - 1)
        return v - (1 << c) if v >> (c - 1) else v

    def ins(e, n, o, c):
        o, c = clamp_oc(o, c)
        if c == 0:
            return e
        mask = ((1 << c) - 1) << o
        return (e & ~mask & 0xffffffff) | ((n << o) & mask)

    F = np.float32

    def ref(p):
        u, i = int(p[:1].view(np.uint32)[0]), int(p[:1].view(np.int32)[0])
        o, c = int(np.trunc(F(abs(p[1]) * F(40)))), int(np.trunc(F(abs(p[2]) * F(40))))
        t = [F(ext_u(u, o, c) & 0xffff), F(F(65536) * F(ext_i(i, o, c) & 0xff)), F(F(ins(u, 0x5a5a5a5a, o, c) >> 20) * F(0.0001)),
             F(F(ext_u(7, 1, 2)) * F(1e8)), F(F(ins(i & 0xffffffff, 0xffffffff, o, c) & 0xfff) * F(1e-8))]
        return F(F(F(F(t[0] + t[1]) + t[2]) + t[3]) + t[4])

    assert f32_equal(va, np.array([ref(p) for p in pts], np.float32)).all()
    assert a.create_shader_module(None).cubin_size > 0


def test_glsl_integer_literal_bit_patterns_and_constant_folding(built):
    """GLSL 4.1.3: an unsuffixed literal with the sign bit set is a negative int (0xFFFFFFFF == -1), more than 32 bits is
    an error; integer constant expressions fold with 32-bit wrap-around like the run-time operations"""
    cases = [("0xFFFFFFFF", -1.0), ("0x80000000", -2147483648.0), ("4294967295", -1.0), ("0xFFFFFFFFu", 4294967296.0), ("2147483647 + 1", -2147483648.0),
             ("1 << 31", -2147483648.0), ("int(uint(-1) >> 1)", 2147483647.0), ("int(3000000000u)", -1294967296.0), ("7 / -2", -3.0), ("-7 % 3", -1.0),
             ("int(-1.5)", -1.0), ("int(1e20)", 2147483647.0), ("int(uint(-5.0))", 0.0), ("65536 * 65536", 0.0), ("abs(-2147483647 - 1)", -2147483648.0),
             ("int(float((1 << 24) + 1))", 16777216.0), ("5 & 3 | 8 ^ 2", 11.0), ("-3 >> 1", -2.0), ("int(uint(-3) >> 1u)", 2147483646.0), ("100 / 10 / 5", 2.0),
             ("1 < 2 == true ? 1 : 0", 1.0)]
    body = "\n".join(f"  if (k == {i}) return float({e});" for i, (e, _) in enumerate(cases))
    dyn = "\n".join(f"  if (k == {100 + i}) return float({e.replace('1 <<', '(1 + z) <<').replace('7 /', '(7 + z) /').replace('-7 %', '(z - 7) %').replace('65536 *', '(65536 + z) *')});"
                    for i, (e, _) in enumerate(cases))
    src = "#version 450 core\nfloat sdf(vec3 p) {\n  int k = int(p.x); int z = int(p.y);\n" + body + "\n" + dyn + "\n  return 0.5;\n}\nvoid main() {}\n"
    sh = s2m.Sdf3DShader.from_source(src, s2m.SRC_GLSL_FRAGMENT, "sdf")
    ks = np.array([i for i in range(len(cases))] + [100 + i for i in range(len(cases))], np.float32)
    pts = np.stack([ks, np.zeros_like(ks), np.zeros_like(ks)], 1)
    got = host_eval.eval_points(sh.lower_to_cuda(), pts)
    want = np.array([w for _, w in cases] * 2, np.float32)
    assert f32_equal(got, want).all(), [(cases[i % len(cases)][0], got[i], want[i]) for i in np.flatnonzero(~f32_equal(got, want))]
    with pytest.raises(s2m.S2mError) as e:
        s2m.Sdf3DShader.from_source("#version 450 core\nfloat sdf(vec3 p) { return float(0x1FFFFFFFF); }\nvoid main() {}\n", s2m.SRC_GLSL_FRAGMENT, "sdf")
    assert e.value.kind == "PARSE" and "32 bits" in str(e.value)


def test_glsl_conditional_evaluates_only_the_chosen_arm(built):
    """`c ? t : f`: an arm that calls a user function is lowered to a temporary + if / else (as naga does), not to
    select() -- the untaken arm's call must not run.  The ADVICE example: select() gave 15 for x <= 0."""
    src = ("float g = 0.; float inc() { g += 1.; return g; }\n"
           "float sdf(vec3 p) { g = 0.; float r = p.x > 0. ? inc() : 5.; return r + 10. * g; }\nvoid main() {}\n")
    sh = s2m.Sdf3DShader.from_source(src, s2m.SRC_GLSL_FRAGMENT, "sdf")
    assert "select(" not in sh.source.split("// sdf3d::normal")[0] and "_cond1 = inc();" in sh.source
    got = host_eval.eval_points(sh.lower_to_cuda(), np.array([[1, 0, 0], [-1, 0, 0]], np.float32))
    assert got.tolist() == [11.0, 5.0]
    # nested conditionals, an assignment inside an arm, an out parameter inside an arm
    src = ("void bump(inout float v) { v += 2.; }\nfloat sq(float x) { return x * x; }\n"
           "float sdf(vec3 p) {\n  float a = 1., b = 0.;\n"
           "  float r = p.z > 0. ? (p.x > 0. ? sq(p.z) : 1.) : 7.;\n"
           "  float s = p.y > 0. ? (b = 3.) : a;\n"
           "  float t = p.x > 1. ? sq(a) : a; if (p.x > 2.) bump(a);\n"
           "  return r + 10. * s + 100. * b + 1000. * t + 10000. * a; }\nvoid main() {}\n")
    sh = s2m.Sdf3DShader.from_source(src, s2m.SRC_GLSL_FRAGMENT, "sdf")
    pts = np.array([[1, 1, 2], [-1, -1, 2], [3, 0, -1]], np.float32)
    want = [4 + 30 + 300 + 1000 + 10000, 1 + 10 + 0 + 1000 + 10000, 7 + 10 + 0 + 1000 + 30000]
    assert host_eval.eval_points(sh.lower_to_cuda(), pts).tolist() == [float(w) for w in want]
    # pure arms stay an expression
    sh = s2m.Sdf3DShader.from_source("float sdf(vec3 p) { return p.x > 0. ? length(p) : -p.y; }\nvoid main() {}\n", s2m.SRC_GLSL_FRAGMENT, "sdf")
    assert "select(" in sh.source


def test_glsl_conditional_with_a_side_effect_where_no_statement_fits_is_refused(built):
    """in an `else if` condition (or right of && / ||) the conditional stays select(); a pure call there is fine,
    a call that assigns module-scope state would run although its arm is not taken -> S2M_ERR_UNSUPPORTED"""
    pure = ("float sq(float x) { return x * x; }\n"
            "float sdf(vec3 p) { float r = 0.; if (p.y > 5.) r = 1.; else if ((p.x > 0. ? sq(p.x) : 5.) > 2.) r = 2.; return r; }\nvoid main() {}\n")
    sh = s2m.Sdf3DShader.from_source(pure, s2m.SRC_GLSL_FRAGMENT, "sdf")
    assert host_eval.eval_points(sh.lower_to_cuda(), np.array([[1, 0, 0], [-1, 0, 0], [3, 0, 0]], np.float32)).tolist() == [0.0, 2.0, 2.0]
    impure = ("float g = 0.; float inc() { g += 1.; return g; }\n"
              "float sdf(vec3 p) { g = 0.; float r = 0.; if (p.y > 5.) r = 1.; else if ((p.x > 0. ? inc() : 5.) > 2.) r = 2.; return r + g; }\nvoid main() {}\n")
    with pytest.raises(s2m.S2mError) as e:
        s2m.Sdf3DShader.from_source(impure, s2m.SRC_GLSL_FRAGMENT, "sdf")
    assert e.value.kind == "UNSUPPORTED" and "?:" in str(e.value)


def test_module_scope_initialiser_that_reads_assigned_state_is_a_validation_error(built):
    """used to surface as a raw compiler error ('G was not declared in this scope')"""
    src = ("var<private> A: f32 = 1.0;\nvar<private> B: f32 = A * 3.0;\nfn bump() { A = A + 1.0; }\n"
           "fn sdf3d(p: vec3f) -> f32 { bump(); let c = vec3(B, A, 1.0); return length(p) - c.x; }\n")
    with pytest.raises(s2m.S2mError) as e:
        s2m.Sdf3DShader.from_source(src).lower_to_cuda()
    assert e.value.kind == "VALIDATION" and "'B'" in str(e.value) and "'A'" in str(e.value)
