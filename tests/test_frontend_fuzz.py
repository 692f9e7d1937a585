"""Differential test of the front-end on generated programs (no GPU).

A seeded generator builds typed expression trees (f32, vec2, vec3) over the exactly-defined part of the
shading languages -- arithmetic, comparisons / select, min max clamp mix step smoothstep, floor ceil
trunc round fract sign abs sqrt, dot length distance normalize cross reflect, constructors and
swizzles, 2x2 / 3x3 matrices (columns, products with matrices, vectors and scalars, sums, transpose, determinant) -- prints every tree as WGSL and as GLSL, and evaluates it in numpy with one IEEE f32
operation per source operation (the semantics DESIGN.md section 3 pins; csrc/s2m_math.h, s2m_vec.h).
The front-end's CUDA C++ for the WGSL text and for the GLSL text (through the GLSL -> naga-IR-shaped
path the reference takes, /root/reference/src/shadertoy.rs:169-194), compiled for the host, must
reproduce the numpy values bit for bit.  host_eval.eval_points also runs the packed (f32x2) form of
each program against the scalar one.
"""
import numpy as np
import pytest

import sdf2mesh_b200 as s2m
from sdf2mesh_b200 import _capi
from tests.support import host_eval

F = np.float32
CONSTS = [0.0, 0.5, 1.0, 1.5, 2.0, 0.25, 3.0, 0.125, -0.75, -1.0, 0.375, 4.0, -2.5]
DIM = {"f": 1, "v2": 2, "v3": 3}


# ------------------------------------------------------------------ pinned semantics in numpy (arrays of f32)
def _bits(a):
    return np.ascontiguousarray(a, F).view(np.uint32)


def s_min(a, b):
    eq = np.where(a == b, (_bits(a) | _bits(b)).view(F), np.where(a < b, a, b))   # equal: -0 wins
    return np.where(np.isnan(a), b, np.where(np.isnan(b), a, eq)).astype(F)


def s_max(a, b):
    eq = np.where(a == b, (_bits(a) & _bits(b)).view(F), np.where(a > b, a, b))   # equal: +0 wins
    return np.where(np.isnan(a), b, np.where(np.isnan(b), a, eq)).astype(F)


def s_clamp(x, lo, hi):
    return s_min(s_max(x, lo), hi)


def s_mix(a, b, t):
    return (a * (F(1) - t) + b * t).astype(F)


def s_sign(x):
    return np.where(x > 0, F(1), np.where(x < 0, F(-1), np.where(x == 0, F(0), x))).astype(F)


def s_smoothstep(lo, hi, x):
    t = s_clamp(((x - lo) / (hi - lo)).astype(F), F(0), F(1))
    return (t * t * (F(3) - F(2) * t)).astype(F)


def s_dot(a, b):
    acc = (a[..., 0] * b[..., 0]).astype(F)
    for k in range(1, a.shape[-1]):
        acc = (acc + (a[..., k] * b[..., k]).astype(F)).astype(F)
    return acc


def s_length(a):
    return np.sqrt(s_dot(a, a)).astype(F)


def s_cross(a, b):
    return np.stack([a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1], a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
                     a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]], axis=-1).astype(F)


def bc(x, like):
    """scalar (N,) against vector (N,k)"""
    return x[..., None] if like.ndim == 2 and x.ndim == 1 else x


# ------------------------------------------------------------------ expression trees
class Node:
    def __init__(self, op, ty, kids=(), arg=None):
        self.op, self.ty, self.kids, self.arg = op, ty, tuple(kids), arg


def gen(rng, ty, depth, env=None):
    """a random tree of type ty ('f' | 'v2' | 'v3' | 'b'); env = {"f": [names], "v3": [names], "trig": bool}
    adds local variables as leaves and (trig) sin / cos of shared arguments"""
    if env is not None:
        return _gen_env(rng, ty, depth, env)
    return _gen(rng, ty, depth)


def _gen_env(rng, ty, depth, env):
    """gen() with env: the recursion of _gen below, with some leaves replaced by variables afterwards"""
    t = _gen(rng, ty, depth)

    def patch(n):
        kids = tuple(patch(c) for c in n.kids)
        if n.op == "comp" and env.get("f") and rng.random() < 0.6:
            return Node("var", "f", (), str(rng.choice(env["f"])))
        if n.op == "qswz" and n.ty == "v3" and env.get("v3") and rng.random() < 0.6:
            return Node("var", "v3", (), str(rng.choice(env["v3"])))
        if n.op == "sqrtabs" and env.get("trig"):   # sin(e) * cos(e): the optimizer pairs them
            return Node("sincos", "f", kids)
        return Node(n.op, n.ty, kids, n.arg)
    return patch(t)


def _gen(rng, ty, depth):
    gen = _gen
    r = rng.random()
    if ty == "b":
        a, b = gen(rng, "f", depth - 1), gen(rng, "f", depth - 1)
        return Node(rng.choice(["<", "<=", ">", ">=", "==", "!="]), "b", (a, b))
    if depth <= 0 or r < 0.12:
        if ty == "f":
            return Node("comp", "f", (), int(rng.integers(3))) if rng.random() < 0.65 else Node("const", "f", (), float(rng.choice(CONSTS)))
        if rng.random() < 0.7:
            idx = tuple(int(i) for i in rng.integers(0, 3, DIM[ty]))
            return Node("qswz", ty, (), idx)
        return Node("ctor", ty, tuple(Node("const", "f", (), float(rng.choice(CONSTS))) for _ in range(DIM[ty])))
    d = depth - 1
    if ty == "f":
        choices = ["+", "-", "*", "/", "neg", "abs", "min", "max", "clamp", "mix", "floor", "ceil", "trunc", "round", "fract", "sign",
                   "step", "sqrtabs", "length", "dot", "distance", "pick", "select", "smoothstep", "fmod", "det", "elem"]
        op = rng.choice(choices)
        if op in ("+", "-", "*", "/", "min", "max", "step", "fmod"):
            return Node(op, "f", (gen(rng, "f", d), gen(rng, "f", d)))
        if op in ("neg", "abs", "floor", "ceil", "trunc", "round", "fract", "sign", "sqrtabs"):
            return Node(op, "f", (gen(rng, "f", d),))
        if op in ("clamp", "mix", "smoothstep"):
            return Node(op, "f", (gen(rng, "f", d), gen(rng, "f", d), gen(rng, "f", d)))
        if op == "length":
            return Node(op, "f", (gen(rng, rng.choice(["v2", "v3"]), d),))
        if op in ("dot", "distance"):
            vt = rng.choice(["v2", "v3"])
            return Node(op, "f", (gen(rng, vt, d), gen(rng, vt, d)))
        if op == "pick":
            vt = rng.choice(["v2", "v3"])
            return Node(op, "f", (gen(rng, vt, d),), int(rng.integers(DIM[vt])))
        if op == "det":
            return Node("det", "f", (gen_mat(rng, int(rng.choice([2, 3])), d),))
        if op == "elem":
            n = int(rng.choice([2, 3]))
            return Node("elem", "f", (gen_mat(rng, n, d),), (int(rng.integers(n)), int(rng.integers(n))))
        return Node("select", "f", (gen(rng, "f", d), gen(rng, "f", d), gen(rng, "b", d)))
    choices = ["+", "-", "*", "/", "vs*", "sv*", "vs/", "vs+", "neg", "abs", "min", "max", "clamp", "mixs", "mixv", "floor", "fract", "sign",
               "step", "normalize", "reflect", "ctor", "swz", "select"]
    if ty == "v3":
        choices += ["cross", "ctor21", "ctor12", "widen"]
    else:
        choices += ["narrow"]
    choices += ["mv", "vm", "col"]
    op = rng.choice(choices)
    if op == "mv":   # matrix * vector
        return Node("mv", ty, (gen_mat(rng, DIM[ty], d), gen(rng, ty, d)))
    if op == "vm":   # vector * matrix
        return Node("vm", ty, (gen(rng, ty, d), gen_mat(rng, DIM[ty], d)))
    if op == "col":
        return Node("col", ty, (gen_mat(rng, DIM[ty], d),), int(rng.integers(DIM[ty])))
    if op in ("+", "-", "*", "/", "min", "max", "step", "reflect", "cross"):
        return Node(op, ty, (gen(rng, ty, d), gen(rng, ty, d)))
    if op in ("vs*", "vs/", "vs+"):
        return Node(op, ty, (gen(rng, ty, d), gen(rng, "f", d)))
    if op == "sv*":
        return Node(op, ty, (gen(rng, "f", d), gen(rng, ty, d)))
    if op in ("neg", "abs", "floor", "fract", "sign", "normalize"):
        return Node(op, ty, (gen(rng, ty, d),))
    if op in ("clamp", "mixv"):
        return Node(op, ty, (gen(rng, ty, d), gen(rng, ty, d), gen(rng, ty, d)))
    if op == "mixs":
        return Node(op, ty, (gen(rng, ty, d), gen(rng, ty, d), gen(rng, "f", d)))
    if op == "ctor":
        return Node(op, ty, tuple(gen(rng, "f", d) for _ in range(DIM[ty])))
    if op == "swz":
        return Node(op, ty, (gen(rng, ty, d),), tuple(int(i) for i in rng.integers(0, DIM[ty], DIM[ty])))
    if op == "select":
        return Node(op, ty, (gen(rng, ty, d), gen(rng, ty, d), gen(rng, "b", d)))
    if op == "ctor21":
        return Node(op, "v3", (gen(rng, "v2", d), gen(rng, "f", d)))
    if op == "ctor12":
        return Node(op, "v3", (gen(rng, "f", d), gen(rng, "v2", d)))
    if op == "widen":   # v2.xyx-style swizzle to a wider vector
        return Node(op, "v3", (gen(rng, "v2", d),), tuple(int(i) for i in rng.integers(0, 2, 3)))
    return Node("narrow", "v2", (gen(rng, "v3", d),), tuple(int(i) for i in rng.integers(0, 3, 2)))


def gen_mat(rng, n, depth):
    """a square matrix of size n (type 'm2' | 'm3'): columns, products, sums, scalar multiples, transposes"""
    ty, vt = f"m{n}", f"v{n}"
    d = depth - 1
    if depth <= 0 or rng.random() < 0.35:
        return Node("mcols", ty, tuple(_gen(rng, vt, max(d, 0)) for _ in range(n)))
    op = rng.choice(["mm", "m+", "m-", "ms", "sm", "transpose", "mcols"])
    if op in ("mm", "m+", "m-"):
        return Node(op, ty, (gen_mat(rng, n, d), gen_mat(rng, n, d)))
    if op == "ms":
        return Node(op, ty, (gen_mat(rng, n, d), _gen(rng, "f", d)))
    if op == "sm":
        return Node(op, ty, (_gen(rng, "f", d), gen_mat(rng, n, d)))
    if op == "transpose":
        return Node(op, ty, (gen_mat(rng, n, d),))
    return Node("mcols", ty, tuple(_gen(rng, vt, d) for _ in range(n)))


XYZ = "xyz"


def lit(v):
    s = repr(float(v))
    return s if ("." in s or "e" in s) else s + ".0"


def show(n, glsl):
    """source text of a tree; fully parenthesised except where the test wants precedence to matter"""
    k = [show(c, glsl) for c in n.kids]
    V = {"v2": "vec2" if glsl else "vec2f", "v3": "vec3" if glsl else "vec3f"}
    op = n.op
    if op == "mcols":
        nn = len(n.kids)
        return (f"mat{nn}(" if glsl else f"mat{nn}x{nn}f(") + ", ".join(k) + ")"
    if op in ("mm", "ms", "sm", "mv", "vm"):
        return f"({k[0]} * {k[1]})"
    if op in ("m+", "m-"):
        return f"({k[0]} {op[1]} {k[1]})"
    if op == "det":
        return f"determinant({k[0]})"
    if op == "elem":
        return f"{k[0]}[{n.arg[0]}].{XYZ[n.arg[1]]}"
    if op == "col":
        return f"{k[0]}[{n.arg}]"
    if op == "comp":
        return f"q.{XYZ[n.arg]}"
    if op == "var":
        return n.arg
    if op == "sincos":
        return f"(sin({k[0]}) * cos({k[0]}))"
    if op == "const":
        return lit(n.arg) if n.arg >= 0 else f"({lit(n.arg)})"
    if op == "qswz":
        return "q." + "".join(XYZ[i] for i in n.arg)
    if op in ("ctor", "ctor21", "ctor12"):
        return f"{V[n.ty]}({', '.join(k)})"
    if op in ("+", "-", "*", "/"):
        return f"({k[0]} {op} {k[1]})"
    if op in ("vs*", "sv*"):
        return f"({k[0]} * {k[1]})"
    if op == "vs/":
        return f"({k[0]} / {k[1]})"
    if op == "vs+":
        return f"({k[0]} + {k[1]})"
    if op in ("<", "<=", ">", ">=", "==", "!="):
        return f"({k[0]} {op} {k[1]})"
    if op == "neg":
        return f"(-{k[0]})"
    if op == "sqrtabs":
        return f"sqrt(abs({k[0]}))"
    if op == "fmod":
        return f"({k[0]} - {k[1]} * trunc({k[0]} / {k[1]}))" if glsl else f"({k[0]} % {k[1]})"
    if op in ("mixs", "mixv"):
        if op == "mixs" and not glsl:
            return f"mix({k[0]}, {k[1]}, {V[n.ty]}({k[2]}))"   # exercise both: WGSL splat, GLSL mix(v, v, float)
        return f"mix({k[0]}, {k[1]}, {k[2]})"
    if op == "pick":
        return f"{k[0]}.{XYZ[n.arg]}"
    if op in ("swz", "widen", "narrow"):
        return f"{k[0]}." + "".join(XYZ[i] for i in n.arg)
    if op == "select":
        return f"({k[2]} ? {k[1]} : {k[0]})" if glsl else f"select({k[0]}, {k[1]}, {k[2]})"
    if op == "round":
        return f"roundEven({k[0]})" if glsl else f"round({k[0]})"
    return f"{op}({', '.join(k)})"


def evaluate(n, q):
    """numpy value of a tree at the points q (N,3): (N,) for f / b, (N,k) for vectors"""
    k = [evaluate(c, q) for c in n.kids]
    op = n.op
    N = q.shape[0]
    with np.errstate(all="ignore"):
        # matrices: (N, column, row)
        if op == "mcols":
            return np.stack(k, axis=1).astype(F)
        if op == "mv" or op == "mm":
            def mat_vec(m, v):   # per row: c0.r * v.x + c1.r * v.y + ..., left to right
                acc = (m[:, 0, :] * v[:, 0:1]).astype(F)
                for c in range(1, m.shape[1]):
                    acc = (acc + (m[:, c, :] * v[:, c:c + 1]).astype(F)).astype(F)
                return acc
            if op == "mv":
                return mat_vec(k[0], k[1])
            return np.stack([mat_vec(k[0], k[1][:, j, :]) for j in range(k[1].shape[1])], axis=1).astype(F)
        if op == "vm":
            return np.stack([s_dot(k[0], k[1][:, c, :]) for c in range(k[1].shape[1])], axis=-1).astype(F)
        if op == "m+":
            return (k[0] + k[1]).astype(F)
        if op == "m-":
            return (k[0] - k[1]).astype(F)
        if op == "ms":
            return (k[0] * k[1][:, None, None]).astype(F)
        if op == "sm":
            return (k[0][:, None, None] * k[1]).astype(F)
        if op == "transpose":
            return np.swapaxes(k[0], 1, 2).copy()
        if op == "det":
            m = k[0]
            if m.shape[1] == 2:
                return ((m[:, 0, 0] * m[:, 1, 1]).astype(F) - (m[:, 1, 0] * m[:, 0, 1]).astype(F)).astype(F)
            return s_dot(m[:, 0, :], s_cross(m[:, 1, :], m[:, 2, :]))
        if op == "elem":
            return k[0][:, n.arg[0], n.arg[1]]
        if op == "col":
            return k[0][:, n.arg, :]
        if op == "comp":
            return q[:, n.arg]
        if op == "const":
            return np.full(N, n.arg, F)
        if op == "qswz":
            return q[:, list(n.arg)]
        if op == "ctor":
            return np.stack(k, axis=-1).astype(F)
        if op == "ctor21":
            return np.concatenate([k[0], k[1][:, None]], axis=-1).astype(F)
        if op == "ctor12":
            return np.concatenate([k[0][:, None], k[1]], axis=-1).astype(F)
        if op in ("+", "vs+"):
            return (k[0] + bc(k[1], k[0])).astype(F)
        if op == "-":
            return (k[0] - k[1]).astype(F)
        if op in ("*", "vs*"):
            return (k[0] * bc(k[1], k[0])).astype(F)
        if op == "sv*":
            return (bc(k[0], k[1]) * k[1]).astype(F)
        if op in ("/", "vs/"):
            return (k[0] / bc(k[1], k[0])).astype(F)
        if op in ("<", "<=", ">", ">=", "==", "!="):
            return {"<": np.less, "<=": np.less_equal, ">": np.greater, ">=": np.greater_equal, "==": np.equal, "!=": np.not_equal}[op](k[0], k[1])
        if op == "neg":
            return (-k[0]).astype(F)
        if op == "abs":
            return np.abs(k[0]).astype(F)
        if op == "min":
            return s_min(k[0], k[1])
        if op == "max":
            return s_max(k[0], k[1])
        if op == "clamp":
            return s_clamp(k[0], k[1], k[2])
        if op in ("mix", "mixv"):
            return s_mix(k[0], k[1], k[2])
        if op == "mixs":
            return s_mix(k[0], k[1], bc(k[2], k[0]))
        if op == "floor":
            return np.floor(k[0]).astype(F)
        if op == "ceil":
            return np.ceil(k[0]).astype(F)
        if op == "trunc":
            return np.trunc(k[0]).astype(F)
        if op == "round":
            return np.rint(k[0]).astype(F)
        if op == "fract":
            return (k[0] - np.floor(k[0])).astype(F)
        if op == "sign":
            return s_sign(k[0])
        if op == "step":
            return np.where(k[0] <= k[1], F(1), F(0)).astype(F)
        if op == "sqrtabs":
            return np.sqrt(np.abs(k[0])).astype(F)
        if op == "smoothstep":
            return s_smoothstep(k[0], k[1], k[2])
        if op == "fmod":
            return (k[0] - (k[1] * np.trunc((k[0] / k[1]).astype(F))).astype(F)).astype(F)
        if op == "length":
            return s_length(k[0])
        if op == "dot":
            return s_dot(k[0], k[1])
        if op == "distance":
            return s_length((k[0] - k[1]).astype(F))
        if op == "normalize":
            return (k[0] / s_length(k[0])[:, None]).astype(F)
        if op == "cross":
            return s_cross(k[0], k[1])
        if op == "reflect":   # i - (2*dot(n,i))*n
            return (k[0] - ((F(2) * s_dot(k[1], k[0])).astype(F)[:, None] * k[1]).astype(F)).astype(F)
        if op == "pick":
            return k[0][:, n.arg]
        if op in ("swz", "widen", "narrow"):
            return k[0][:, list(n.arg)]
        if op == "select":
            c = k[2] if k[0].ndim == 1 else k[2][:, None]
            return np.where(c, k[1], k[0]).astype(F)
    raise AssertionError(op)


def program(trees, glsl):
    fns = []
    for i, t in enumerate(trees):
        body = show(t, glsl)
        fns.append(f"float g{i}(vec3 q) {{ return {body}; }}" if glsl else f"fn g{i}(q: vec3f) -> f32 {{ return {body}; }}")
    if glsl:
        sel = "\n".join(f"    if (k == {i}) return g{i}(q);" for i in range(len(trees)))
        main = ("float sdf(vec3 p) {\n    int k = int(floor(p.z));\n    vec3 q = vec3(p.x, p.y, p.z - float(k));\n" + sel +
                "\n    return 0.0;\n}\nvoid main() {}\n")
        return "#version 450 core\n" + "\n".join(fns) + "\n" + main
    sel = "\n".join(f"        case {i}: {{ return g{i}(q); }}" for i in range(len(trees)))
    main = ("fn sdf3d(p: vec3f) -> f32 {\n    let k = i32(floor(p.z));\n    let q = vec3f(p.x, p.y, p.z - f32(k));\n    switch k {\n" + sel +
            "\n        default: { return 0.0; }\n    }\n}\n")
    return "\n".join(fns) + "\n" + main


N_FUNCS = 48
PTS_PER_FUNC = 160


def sample_points(rng, n_funcs):
    """PTS_PER_FUNC points per function: selector k in the integer part of z; q spans magnitudes and signs, with
    exact ties (multiples of 0.5) mixed in so that min/max/step/round/select see equal operands"""
    n = n_funcs * PTS_PER_FUNC
    q = rng.uniform(-2.0, 2.0, (n, 3)).astype(F)
    tie = rng.random((n, 3)) < 0.25
    q[tie] = (np.round(q[tie] * 2) / 2).astype(F)
    q[:, 2] = np.floor((np.abs(q[:, 2]) % F(1.0)) * F(4096)) / F(4096)   # fractional part carried by z: a multiple of 2^-12 below 1, so that z + k is exact
    k = np.repeat(np.arange(n_funcs), PTS_PER_FUNC)
    p = q.copy()
    p[:, 2] = (q[:, 2] + k.astype(F)).astype(F)
    q[:, 2] = (p[:, 2] - k.astype(F)).astype(F)   # what the program recomputes (exact for these magnitudes)
    return p, q, k


def same(a, b):
    return (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))


@pytest.mark.parametrize("seed", [11, 12, 13, 14])
def test_generated_programs_wgsl_and_glsl_equal_numpy(built, seed):
    rng = np.random.default_rng(seed)
    trees = [gen(rng, "f", int(rng.integers(2, 6))) for _ in range(N_FUNCS)]
    p, q, k = sample_points(rng, N_FUNCS)
    want = np.zeros(p.shape[0], F)
    for i, t in enumerate(trees):
        m = k == i
        want[m] = evaluate(t, q[m])
    assert np.isfinite(want).mean() > 0.5   # the programs are not all-NaN
    for glsl in (False, True):
        text = program(trees, glsl)
        sh = s2m.Sdf3DShader.from_source(text, _capi.SRC_GLSL_FRAGMENT if glsl else _capi.SRC_SDF3D, "sdf")
        body = sh.lower_to_cuda()
        host_eval.register_packed(body, sh.lower_to_cuda_packed())
        got = host_eval.eval_points(body, p)
        ok = same(got, want)
        if not ok.all():
            bad = int(np.flatnonzero(~ok)[0])
            i = int(k[bad])
            raise AssertionError(f"{'GLSL' if glsl else 'WGSL'} seed {seed} g{i}: {show(trees[i], glsl)}\n at q = {q[bad].tolist()}: "
                                 f"got {got[bad]!r}, numpy {want[bad]!r} ({int((~ok).sum())} mismatches in all)")


def test_generated_program_compiles_for_sm100a(built):
    """one generated program through NVRTC (no device needed): the emitter's output is valid device code too"""
    rng = np.random.default_rng(5)
    trees = [gen(rng, "f", 4) for _ in range(12)]
    for glsl in (False, True):
        sh = s2m.Sdf3DShader.from_source(program(trees, glsl), _capi.SRC_GLSL_FRAGMENT if glsl else _capi.SRC_SDF3D, "sdf")
        assert sh.create_shader_module(None).cubin_size > 0


# ------------------------------------------------------------------ statements: the IR optimizer against itself
def gen_loop_function(rng, i):
    """GLSL: locals, a counted loop whose body is `prefix; if (c) continue|break; rest` with assignments,
    compound assignments, if/else and sin/cos pairs -- the shapes frontend/optimize.cpp rewrites
    (continue -> break, sin/cos pairing) next to shapes it must leave alone (prefix reads what it
    writes, condition or prefix uses the counter, the counter lives on after the loop)"""
    env = {"f": ["a", "b", "acc"], "v3": ["z", "q"], "trig": True}
    pure = {"f": ["b"], "v3": ["z", "q"], "trig": True}   # what an idempotent prefix writing `a` may read

    def e(ty, depth, en=env):
        return show(gen(rng, ty, depth, en), True)

    def assign(en=env):
        r = rng.random()
        if r < 0.3:
            return f"a = {e('f', 2, en)};"
        if r < 0.45:
            return f"b {rng.choice(['=', '+=', '*=', '-='])} {e('f', 2, en)};"
        if r < 0.6:
            return f"acc += {e('f', 2, en)};"
        if r < 0.8:
            return f"z = {e('v3', 2, en)};"
        if r < 0.9:
            return f"z.{rng.choice(['x', 'y', 'z'])} = {e('f', 2, en)};"
        return f"if ({e('b', 2, en)}) {{ a = {e('f', 1, en)}; }} else {{ z = {e('v3', 1, en)}; acc += 0.5; }}"

    n_iter = int(rng.integers(2, 7))
    shape = rng.random()
    uses_counter = shape > 0.8
    outer_counter = 0.7 < shape <= 0.8
    if shape < 0.55:   # the rewritable shape: a = f(b, z, q); if (c(a, b, z)) continue; rest changes b, z, acc
        prefix = [f"a = {e('f', 3, pure)};"] + ([f"float t = {e('f', 2, pure)};", "a = a * 0.5 + t;"] if rng.random() < 0.4 else [])
        cond = show(gen(rng, "b", 2, {"f": ["a", "b"], "v3": ["z"], "trig": False}), True)
    else:
        prefix = [assign() for _ in range(int(rng.integers(1, 4)))]
        cond = e("b", 2)
    if uses_counter:
        prefix.append("a += float(i) * 0.25;")
    jump = "continue" if rng.random() < 0.8 else "break"
    rest = [assign() for _ in range(int(rng.integers(1, 4)))]
    if rng.random() < 0.3:
        rest.insert(int(rng.integers(0, len(rest) + 1)), f"if ({e('b', 1)}) continue;")
    body = "\n        ".join(prefix + [f"if ({cond}) {jump};"] + rest)
    head = f"int i = 0;\n    for (; i < {n_iter}; i++)" if outer_counter else f"for (int i = 0; i < {n_iter}; i++)"
    tail = " + float(i)" if outer_counter else ""
    return (f"float g{i}(vec3 q) {{\n    float a = {e('f', 2, {'f': [], 'v3': ['q'], 'trig': False})};\n    float b = q.y * 0.5;\n    float acc = 0.0;\n    vec3 z = q;\n"
            f"    {head} {{\n        {body}\n    }}\n    return a + b + acc + dot(z, vec3(1.0, 0.5, 0.25)){tail};\n}}")


@pytest.mark.parametrize("seed", [21, 22, 23])
def test_generated_loops_optimized_equals_unoptimized(built, monkeypatch, seed):
    rng = np.random.default_rng(seed)
    n = 40
    fns = [gen_loop_function(rng, i) for i in range(n)]
    sel = "\n".join(f"    if (k == {i}) return g{i}(q);" for i in range(n))
    text = ("#version 450 core\n" + "\n".join(fns) + "\nfloat sdf(vec3 p) {\n    int k = int(floor(p.z));\n    vec3 q = vec3(p.x, p.y, p.z - float(k));\n" +
            sel + "\n    return 0.0;\n}\nvoid main() {}\n")
    p, _, _ = sample_points(rng, n)
    monkeypatch.delenv("S2M_NO_IR_OPT", raising=False)
    sh = s2m.Sdf3DShader.from_source(text, _capi.SRC_GLSL_FRAGMENT, "sdf")
    opt = sh.lower_to_cuda()
    host_eval.register_packed(opt, sh.lower_to_cuda_packed())
    monkeypatch.setenv("S2M_NO_IR_OPT", "1")
    plain = s2m.Sdf3DShader.from_source(text, _capi.SRC_GLSL_FRAGMENT, "sdf").lower_to_cuda()
    monkeypatch.delenv("S2M_NO_IR_OPT", raising=False)
    # the optimizer did something on this program, and not everywhere
    n_loops = plain.count("for (") + plain.count("while (")
    # (a rotated loop -- optimize.cpp rotate_guarded_loop -- carries one more `break` than the loop it came from)
    assert 0 < opt.count("break;") - opt.count("for (; ; )") - plain.count("break;") < n_loops
    assert opt.count("for (; ; )") > 0
    assert opt.count("f_sincos_pair(") > 0 and "f_sincos_pair(" not in plain
    a, b = host_eval.eval_points(opt, p), host_eval.eval_points(plain, p)
    ok = same(a, b)
    assert np.isfinite(b).mean() > 0.3
    if not ok.all():
        bad = int(np.flatnonzero(~ok)[0])
        raise AssertionError(f"seed {seed}: optimized {a[bad]!r} != unoptimized {b[bad]!r} at p = {p[bad].tolist()} in\n{fns[bad // PTS_PER_FUNC]}")


# ------------------------------------------------------------------ integers: i32 / u32 trees
# Semantics pinned in csrc/s2m_vec.h (WGSL's): two's-complement wrap-around, x / 0 = x, x % 0 = 0,
# INT_MIN / -1 = INT_MIN, INT_MIN % -1 = 0, truncating division, shift counts taken mod 32, arithmetic >> on
# i32, float -> int conversions truncate and saturate (NaN -> 0), int -> float rounds to nearest even.
I64 = np.int64


def wrap_i(x):
    return ((x.astype(I64) + (1 << 31)) % (1 << 32)) - (1 << 31)


def wrap_u(x):
    return x.astype(I64) % (1 << 32)


def gen_int(rng, ty, depth):
    """ty: 'i' (i32) | 'u' (u32)"""
    if depth <= 0 or rng.random() < 0.15:
        r = rng.random()
        if r < 0.45:   # from the point: trunc(q.c * scale), saturating for the big scales
            return Node("fromf", ty, (), (int(rng.integers(3)), float(rng.choice([8.0, 1000.0, 65536.0, 3.0e9, -3.0e9 if ty == "i" else 5.0e9]))))
        if r < 0.6:
            return Node("bits", ty, (), int(rng.integers(3)))   # bitcast of q.c
        c = int(rng.choice([0, 1, 2, 3, 7, 16, 255, 65535, 1664525, 1013904223, 2147483647] + ([-1, -2147483648, -7] if ty == "i" else [4294967295, 2654435769])))
        return Node("iconst", ty, (), c)
    d = depth - 1
    ops = (["+", "-", "*", "/", "%", "&", "|", "^", "<<", ">>", "not", "min", "max", "clamp", "cast", "sel", "popc", "rev", "clz", "ctz", "flb", "ftb"] +
           (["neg", "abs", "sign"] if ty == "i" else []))
    op = rng.choice(ops)
    if op in ("+", "-", "*", "/", "%", "&", "|", "^", "min", "max"):
        return Node(op, ty, (gen_int(rng, ty, d), gen_int(rng, ty, d)))
    if op in ("<<", ">>"):
        cnt = Node("iconst", "u", (), int(rng.integers(0, 32))) if rng.random() < 0.6 else Node("&", "u", (gen_int(rng, "u", d), Node("iconst", "u", (), 31)))
        return Node(op, ty, (gen_int(rng, ty, d), cnt))
    if op in ("not", "neg", "abs", "sign", "popc", "rev", "clz", "ctz", "flb", "ftb"):
        return Node(op, ty, (gen_int(rng, ty, d),))
    if op == "clamp":
        return Node(op, ty, (gen_int(rng, ty, d), gen_int(rng, ty, d), gen_int(rng, ty, d)))
    if op == "cast":   # i32(u) / u32(i): same bits
        return Node("cast", ty, (gen_int(rng, "u" if ty == "i" else "i", d),))
    return Node("sel", ty, (gen_int(rng, ty, d), gen_int(rng, ty, d), Node(rng.choice(["<", ">=", "=="]), "b", (gen_int(rng, ty, d), gen_int(rng, ty, d)))))


def show_int(n, glsl):
    k = [show_int(c, glsl) for c in n.kids]
    T = {"i": "int" if glsl else "i32", "u": "uint" if glsl else "u32"}
    op = n.op
    if op == "fromf":
        return f"{T[n.ty]}(q.{XYZ[n.arg[0]]} * {lit(n.arg[1]) if n.arg[1] >= 0 else '(' + lit(n.arg[1]) + ')'})"
    if op == "bits":
        return (f"floatBitsTo{'Int' if n.ty == 'i' else 'Uint'}(q.{XYZ[n.arg]})" if glsl else f"bitcast<{T[n.ty]}>(q.{XYZ[n.arg]})")
    if op == "iconst":
        if n.ty == "u":
            return f"{n.arg}u"
        if n.arg == -2147483648:
            return "(-2147483647 - 1)" if glsl else "i32(-2147483647 - 1)"   # WGSL: concrete, or -(...) of it would be the abstract-int +2147483648
        s = str(n.arg) if glsl else f"{n.arg}i"
        return f"({s})" if n.arg < 0 else s
    if op in ("+", "-", "*", "/", "%", "&", "|", "^", "<<", ">>", "<", ">=", "=="):
        return f"({k[0]} {op} {k[1]})"
    if op == "not":
        return f"(~{k[0]})"
    if op == "neg":
        return f"(-{k[0]})"
    if op == "cast":
        return f"{T[n.ty]}({k[0]})"
    if op == "sel":
        return f"({k[2]} ? {k[1]} : {k[0]})" if glsl else f"select({k[0]}, {k[1]}, {k[2]})"
    if op in ("popc", "rev", "clz", "ctz", "flb", "ftb"):
        if not glsl:
            return {"popc": "countOneBits", "rev": "reverseBits", "clz": "countLeadingZeros", "ctz": "countTrailingZeros",
                    "flb": "firstLeadingBit", "ftb": "firstTrailingBit"}[op] + f"({k[0]})"
        # GLSL: bitCount / findMSB / findLSB return int whatever the argument; there is no clz / ctz
        back = (lambda e: f"uint({e})") if n.ty == "u" else (lambda e: e)
        if op == "rev":
            return f"bitfieldReverse({k[0]})"
        if op == "popc":
            return back(f"bitCount({k[0]})")
        if op == "flb":
            return back(f"findMSB({k[0]})")
        if op == "ftb":
            return back(f"findLSB({k[0]})")
        if op == "clz":
            return back(f"(31 - findMSB(uint({k[0]})))")
        return back(f"(findLSB({k[0]}) < 0 ? 32 : findLSB({k[0]}))")
    return f"{op}({', '.join(k)})"


def _bit_op(op, ty, v):
    """one value (python int holding the i32 / u32 value)"""
    u = v % (1 << 32)
    if op == "popc":
        return bin(u).count("1")
    if op == "rev":
        r = int(format(u, "032b")[::-1], 2)
        return r - (1 << 32) if ty == "i" and r >= (1 << 31) else r
    if op == "clz":
        return 32 - u.bit_length()
    if op == "ctz":
        return 32 if u == 0 else (u & -u).bit_length() - 1
    if op == "ftb":
        return (-1 if ty == "i" else (1 << 32) - 1) if u == 0 else (u & -u).bit_length() - 1
    if ty == "u":   # flb
        return (1 << 32) - 1 if u == 0 else u.bit_length() - 1
    t = ~v if v < 0 else v
    return -1 if t == 0 else t.bit_length() - 1


def eval_int(n, q):
    """int64 arrays holding the i32 / u32 value"""
    k = [eval_int(c, q) for c in n.kids]
    w = wrap_i if n.ty == "i" else wrap_u
    op = n.op
    N = q.shape[0]
    if op == "fromf":
        with np.errstate(all="ignore"):
            f = (q[:, n.arg[0]] * F(n.arg[1])).astype(F).astype(np.float64)
        lo, hi = (-(1 << 31), (1 << 31) - 1) if n.ty == "i" else (0, (1 << 32) - 1)
        t = np.where(np.isnan(f), 0.0, np.clip(np.trunc(f), lo, hi))
        return t.astype(I64)
    if op == "bits":
        b = np.ascontiguousarray(q[:, n.arg]).view(np.uint32).astype(I64)
        return wrap_i(b) if n.ty == "i" else b
    if op == "iconst":
        return np.full(N, n.arg, I64)
    if op == "+":
        return w(k[0] + k[1])
    if op == "-":
        return w(k[0] - k[1])
    if op == "*":
        a, b = k[0].astype(object), k[1].astype(object)   # exact products, then wrap
        return w(np.array([int(x) * int(y) % (1 << 64) for x, y in zip(a, b)], dtype=np.uint64).astype(I64))
    if op in ("/", "%"):
        a, b = k[0], k[1]
        bad = (b == 0) | ((a == -(1 << 31)) & (b == -1) & (n.ty == "i"))
        bs = np.where(bad, 1, b)
        quo = np.sign(a) * np.sign(bs) * (np.abs(a) // np.abs(bs))
        if op == "/":
            return np.where(bad, a, quo)
        return np.where(bad, 0, a - quo * bs)
    if op == "&":
        return w(wrap_u(k[0]) & wrap_u(k[1]))
    if op == "|":
        return w(wrap_u(k[0]) | wrap_u(k[1]))
    if op == "^":
        return w(wrap_u(k[0]) ^ wrap_u(k[1]))
    if op == "<<":
        return w(wrap_u(k[0]) << (k[1] % 32))
    if op == ">>":
        return k[0] >> (k[1] % 32)   # arithmetic on the signed value, logical on the unsigned one
    if op in ("popc", "rev", "clz", "ctz", "flb", "ftb"):
        return np.array([_bit_op(op, n.ty, int(v)) for v in k[0]], I64)
    if op == "not":
        return w(~k[0])
    if op == "neg":
        return w(-k[0])
    if op == "abs":
        return w(np.abs(k[0]))
    if op == "sign":
        return np.sign(k[0])
    if op == "min":
        return np.minimum(k[0], k[1])
    if op == "max":
        return np.maximum(k[0], k[1])
    if op == "clamp":
        return np.minimum(np.maximum(k[0], k[1]), k[2])
    if op == "cast":
        return w(k[0])
    if op == "sel":
        return np.where(k[2], k[1], k[0])
    if op == "<":
        return k[0] < k[1]
    if op == ">=":
        return k[0] >= k[1]
    if op == "==":
        return k[0] == k[1]
    raise AssertionError(op)


@pytest.mark.parametrize("seed", [31, 32, 33])
def test_generated_integer_programs_equal_numpy(built, seed):
    rng = np.random.default_rng(seed)
    trees = [gen_int(rng, str(rng.choice(["i", "u"])), int(rng.integers(2, 6))) for _ in range(N_FUNCS)]
    low_bits = rng.random(N_FUNCS) < 0.5   # f32(e) keeps the top 24 bits; f32(e & 1023) the bottom ones
    p, q, k = sample_points(rng, N_FUNCS)
    want = np.zeros(p.shape[0], F)
    for i, t in enumerate(trees):
        m = k == i
        v = eval_int(t, q[m])
        want[m] = ((wrap_u(v) & 1023) if low_bits[i] else v).astype(F)
    for glsl in (False, True):
        fns = []
        for i, t in enumerate(trees):
            e = show_int(t, glsl)
            e = f"({e} & 1023{'u' if t.ty == 'u' else ''})" if low_bits[i] else e
            fns.append(f"float g{i}(vec3 q) {{ return float({e}); }}" if glsl else f"fn g{i}(q: vec3f) -> f32 {{ return f32({e}); }}")
        text = program(trees, glsl)
        text = text[text.index("float sdf(vec3 p)" if glsl else "fn sdf3d("):]
        text = ("#version 450 core\n" if glsl else "") + "\n".join(fns) + "\n" + text
        sh = s2m.Sdf3DShader.from_source(text, _capi.SRC_GLSL_FRAGMENT if glsl else _capi.SRC_SDF3D, "sdf")
        body = sh.lower_to_cuda()
        host_eval.register_packed(body, sh.lower_to_cuda_packed())
        got = host_eval.eval_points(body, p)
        ok = same(got, want)
        if not ok.all():
            bad = int(np.flatnonzero(~ok)[0])
            i = int(k[bad])
            raise AssertionError(f"{'GLSL' if glsl else 'WGSL'} seed {seed} g{i} ({trees[i].ty}32{', low bits' if low_bits[i] else ''}): {show_int(trees[i], glsl)}\n"
                                 f" at q = {q[bad].tolist()}: got {got[bad]!r}, numpy {want[bad]!r} ({int((~ok).sum())} mismatches in all)")


# ------------------------------------------------------------------ damaged inputs: errors, never crashes
MUTATION_TOKENS = ["(", ")", "{", "}", "[", "]", ";", ",", ".", "+", "-", "*", "/", "%", "=", "==", "<", ">", "<<", ">>", "&", "|", "^", "!", "~", "?", ":", "#", "\\", "\n",
                   "if", "else", "for", "while", "return", "float", "vec3", "fn", "let", "var", "0", "1.0", "1e40", "0x", "1u", "struct", "switch", "case", "default", "break",
                   "continue", "#define", "#if", "#endif", "#else", "/*", "*/", "//", "mat3", "array", "->", "@", "&&", "||", "++", "--", "+=", "vec3f", "f32", "i32", "ptr",
                   "loop", "continuing", "\x00", "\xff", "é"]


def test_damaged_sources_give_errors_not_crashes(built):
    """every example and fixture with random deletions, insertions, duplications and truncations: the library
    answers with an S2mError or a translation (which NVRTC then accepts, packed form included) -- the C++
    front-end runs inside the caller's process, so anything else would take the caller down"""
    import os
    import random
    from tests.conftest import ROOT
    rnd = random.Random(2024)
    corpus = [(open(os.path.join(ROOT, "examples", f)).read(), _capi.SRC_SDF3D) for f in ("torus.sdf3d", "martin_cube.sdf3d", "p_key.sdf3d")]
    corpus.append((open(os.path.join(ROOT, "examples", "mandelmesh.frag")).read(), _capi.SRC_GLSL_FRAGMENT))
    corpus.append((open(os.path.join(ROOT, "tests", "data", "wgsl_features.sdf3d")).read(), _capi.SRC_SDF3D))
    for f in sorted(os.listdir(os.path.join(ROOT, "tests", "data"))):
        if f.endswith(".glsl"):
            corpus.append(("#version 450 core\nuniform float iTime; uniform vec3 iResolution; uniform int iFrame; uniform vec4 iMouse;\n" +
                           open(os.path.join(ROOT, "tests", "data", f)).read() + "\nvoid main() {}\n", _capi.SRC_GLSL_FRAGMENT))
    for text, kind in corpus:   # the undamaged corpus translates
        s2m.Sdf3DShader.from_source(text, kind, "sdf" if "float sdf(" in text else "map").lower_to_cuda()
    accepted = rejected = compiled = 0
    for it in range(500):
        src, kind = rnd.choice(corpus)
        s = src
        for _ in range(rnd.randint(1, 4)):
            m, i = rnd.random(), rnd.randrange(len(s) + 1)
            if m < 0.3:
                s = s[:i] + s[min(len(s), i + rnd.randint(1, 12)):]
            elif m < 0.6:
                s = s[:i] + rnd.choice(MUTATION_TOKENS) + s[i:]
            elif m < 0.75:
                j = min(len(s), i + rnd.randint(1, 30))
                s = s[:i] + s[i:j] * 2 + s[j:]
            elif m < 0.9:
                s = s[:i] + " " + rnd.choice(MUTATION_TOKENS) + " " + s[i:]
            else:
                s = s[:i]
        try:
            sh = s2m.Sdf3DShader.from_source(s, kind, "sdf")
            sh.lower_to_cuda()
            sh.lower_to_cuda_packed()
        except s2m.S2mError as e:
            assert e.kind in ("PARSE", "VALIDATION", "UNSUPPORTED", "MISSING_SDF", "SHADER"), str(e)
            rejected += 1
            continue
        accepted += 1
        if compiled < 12:   # what the front-end lets through must be valid CUDA C++ (NVRTC, no device needed)
            assert sh.create_shader_module(None).cubin_size > 0
            compiled += 1
    assert accepted > 10 and rejected > 300
