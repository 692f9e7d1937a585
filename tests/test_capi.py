"""The C-ABI library loads on a machine without a GPU, exports every symbol include/*.h declares,
fails loudly (no CPU fallback) when asked for a device that is not there, and can still JIT every
example SDF to an sm_100a cubin."""
import ctypes
import os
import re
import subprocess

import pytest

import sdf2mesh_b200 as s2m
from sdf2mesh_b200 import _capi
from tests.conftest import EXAMPLE_INPUTS, ROOT, load_example_shader


def header_functions():
    text = open(os.path.join(ROOT, "include", "sdf2mesh_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(s2m_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(built):
    names = header_functions()
    assert len(names) >= 30
    lib = ctypes.CDLL(_capi.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/sdf2mesh_b200.h but not exported"
        assert n in _capi.SYMBOLS, f"{n} has no ctypes prototype in sdf2mesh_b200/_capi.py"
    assert sorted(_capi.SYMBOLS) == names, "ctypes table and header disagree"


def test_struct_layouts_match_header(built):
    """compile a tiny C program against the header and compare sizeof/offsetof with ctypes"""
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "sdf2mesh_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu\n", sizeof(s2m_mesh_params), offsetof(s2m_mesh_params, dims), offsetof(s2m_mesh_params, slab_budget_bytes),
         sizeof(s2m_timings), sizeof(s2m_result_info), offsetof(s2m_result_info, timings));
  return 0;
}'''
    import tempfile
    d = tempfile.mkdtemp()
    open(os.path.join(d, "t.c"), "w").write(src)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")])
    got = [int(x) for x in subprocess.check_output([os.path.join(d, "t")]).split()]
    P, T, R = _capi.MeshParams, _capi.Timings, _capi.ResultInfo
    assert got == [ctypes.sizeof(P), P.dims.offset, P.slab_budget_bytes.offset, ctypes.sizeof(T), ctypes.sizeof(R), R.timings.offset]


def test_no_gpu_means_loud_failure(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(s2m.S2mError) as e:
        s2m.Context(0)
    assert e.value.kind == "NO_DEVICE" and "no CPU fallback" in str(e.value)


@pytest.mark.parametrize("name", sorted(EXAMPLE_INPUTS))
def test_examples_jit_to_sm100a_cubin(built, name, tmp_path):
    m = load_example_shader(name).create_shader_module(None)  # NVRTC needs no device
    assert m.cubin_size > 10000
    data, size = ctypes.c_void_p(), ctypes.c_size_t()
    _capi.check(_capi.lib().s2m_module_cubin(m._h, ctypes.byref(data), ctypes.byref(size)))
    parts = m.cubins()
    assert len(parts) == 3 and parts[0] == ctypes.string_at(data, size.value)   # three NVRTC programs; part 0 = K1
    for i, (part, kernels) in enumerate(zip(parts, (["s2m_k1_slab"], ["s2m_k4_vertices"], ["s2m_k_eval", "s2m_k_cost_probe"]))):
        cubin = tmp_path / f"m{i}.cubin"
        cubin.write_bytes(part)
        out = subprocess.run(["cuobjdump", "-elf", str(cubin)], capture_output=True, text=True).stdout
        assert "sm_100" in out or "SM100" in out.upper() or "EF_CUDA_SM100" in out
        for k in ("s2m_k1_slab", "s2m_k4_vertices", "s2m_k_eval", "s2m_k_cost_probe"):
            assert (k + "\n" in out or k + " " in out or ".text." + k in out) == (k in kernels), (i, k)


def test_jit_split_off_is_one_program(built, monkeypatch, tmp_path):
    """S2M_JIT_SPLIT=0: one NVRTC program holding every kernel (what an offline nvcc build of the unit gives)"""
    monkeypatch.setenv("S2M_JIT_SPLIT", "0")
    m = load_example_shader("torus").create_shader_module(None)
    parts = m.cubins()
    assert len(parts) == 1 and m.cubin_size == len(parts[0])
    cubin = tmp_path / "one.cubin"
    cubin.write_bytes(parts[0])
    out = subprocess.run(["cuobjdump", "-elf", str(cubin)], capture_output=True, text=True).stdout
    for k in ("s2m_k1_slab", "s2m_k4_vertices", "s2m_k_eval", "s2m_k_cost_probe", "s2m_k_eval2"):
        assert k in out


def test_params_from_cli(built):
    """AppState::from(&Arguments): defaults 256 / 2.0, power-of-two rounding UP (main.rs:141-150)"""
    p, r = s2m.params_from_cli(None, None)
    assert list(p.dims) == [256, 256, 256] and not r and list(p.bb_min) == [-1, -1, -1] and list(p.bb_max) == [1, 1, 1]
    assert abs(p.eps - 1e-4) < 1e-10
    for res, want in [(128, 128), (100, 128), (129, 256), (3, 4), (2048, 2048), (2049, 4096)]:
        p, r = s2m.params_from_cli(res, 5.0)
        assert p.dims[0] == want and r == (res != want) and p.bb_max[2] == 2.5


def test_cubin_cache(built, tmp_path, monkeypatch):
    """S2M_CACHE_DIR: the second compile of the same SDF loads the cubin from disk; another SDF gets another file"""
    import sdf2mesh_b200 as s2m
    monkeypatch.setenv("S2M_CACHE_DIR", str(tmp_path))
    src = "fn sdf3d(p: vec3f) -> f32 { return length(p) - 0.75; }"
    a = s2m.Sdf3DShader.from_source(src).create_shader_module(None)
    assert "loaded from" not in a.log and len(list(tmp_path.glob("*.cubin"))) == 3   # one file per NVRTC program
    b = s2m.Sdf3DShader.from_source(src).create_shader_module(None)
    assert b.log.count("cubin loaded from") == 3 and b.cubins() == a.cubins()
    s2m.Sdf3DShader.from_source(src.replace("0.75", "0.5")).create_shader_module(None)
    assert len(list(tmp_path.glob("*.cubin"))) == 6 and not list(tmp_path.glob("*.tmp*"))
    # a damaged entry is ignored and rewritten
    f = sorted(tmp_path.glob("*.cubin"))[0]
    f.write_bytes(b"garbage" * 20)
    monkeypatch.setenv("S2M_CACHE_DIR", str(tmp_path))
    for s in (src, src.replace("0.75", "0.5")):
        assert s2m.Sdf3DShader.from_source(s).create_shader_module(None).cubin_size > 1000
    assert f.read_bytes()[:4] == b"\x7fELF"


def test_packed_k1_policy_and_fallback(built, monkeypatch):
    """K1 in packed f32x2 arithmetic (csrc/s2m_pvec.h) is a per-module decision: default policy, the
    S2M_K1_PACKED override, shaders without a packed form, and -- if NVRTC rejects the packed translation
    unit -- a second compile without it instead of an error."""
    from tests.conftest import load_example_shader
    monkeypatch.delenv("S2M_K1_PACKED", raising=False)
    assert load_example_shader("mandelbulb").create_shader_module(None).packed      # 7 transcendental calls
    assert load_example_shader("torus").create_shader_module(None).packed           # tiny
    assert not load_example_shader("p_key").create_shader_module(None).packed
    monkeypatch.setenv("S2M_K1_PACKED", "0")
    assert not load_example_shader("mandelbulb").create_shader_module(None).packed
    monkeypatch.setenv("S2M_K1_PACKED", "1")
    m = load_example_shader("p_key").create_shader_module(None)
    assert m.packed and "namespace s2m_user_p" in m.cuda_source and "S2M_PACKED_SQRT" not in m.cuda_source
    mat = s2m.Sdf3DShader.from_source("fn sdf3d(p: vec3f) -> f32 { let m = mat2x2f(0.0, 1.0, -1.0, 0.0); return length(m * p.xy) - 1.0; }")
    assert mat.lower_to_cuda_packed() == "" and not mat.create_shader_module(None).packed   # matrices: no packed form
    monkeypatch.setenv("S2M_TEST_BREAK_PACKED", "1")
    m = load_example_shader("torus").create_shader_module(None)
    assert not m.packed and m.cubin_size > 0
    assert "packed (f32x2) form rejected" in m.log and "S2M_TEST_BREAK_PACKED" in m.log


def test_rust_binding_declares_the_same_abi():
    """bindings/rust/ cannot be compiled here (no rustc); at least its extern block must name exactly the
    header's functions, and its constants must carry the header's values"""
    ffi = open(os.path.join(ROOT, "bindings", "rust", "sdf2mesh-b200", "src", "ffi.rs")).read()
    assert sorted(set(re.findall(r"pub fn (s2m_[a-z0-9_]+)\(", ffi))) == header_functions()
    hdr = open(os.path.join(ROOT, "include", "sdf2mesh_b200.h")).read()
    consts = dict(re.findall(r"pub const (S2M_[A-Z0-9_]+): (?:c_int|u32) = (\d+);", ffi))
    assert len(consts) >= 25
    for name, value in consts.items():
        m = re.search(r"\b%s\s*=\s*(\d+)" % name, hdr) or re.search(r"#define\s+%s\s+(\d+)u" % name, hdr)
        assert m, f"{name} is not in the header"
        assert m.group(1) == value, f"{name}: Rust says {value}, header says {m.group(1)}"
    # struct fields in header order
    for struct in ("s2m_mesh_params", "s2m_timings", "s2m_result_info"):
        body_c = re.search(r"typedef struct %s \{(.*?)\} %s;" % (struct, struct), hdr, re.S).group(1)
        body_c = re.sub(r"/\*.*?\*/", "", body_c, flags=re.S)
        fields_c = [f for decl in body_c.split(";") for f in re.findall(r"\*?\s*([a-z_0-9]+)(?:\[\d+\])?\s*(?:,|$)", decl.strip().split(" ", 1)[-1] if decl.strip() else "")]
        body_r = re.search(r"pub struct %s \{(.*?)\n\}" % struct, ffi, re.S).group(1)
        fields_r = re.findall(r"pub ([a-z_0-9]+):", body_r)
        assert fields_r == [f for f in fields_c if f], f"{struct}: {fields_r} vs {fields_c}"


def test_missing_library_fails_loudly(built, tmp_path):
    """no CPU fallback: without the built shared library the package refuses to do anything (a copy of the package
    without its .so, in a fresh interpreter)"""
    import shutil
    import sys
    pkg = tmp_path / "sdf2mesh_b200"
    shutil.copytree(os.path.join(ROOT, "sdf2mesh_b200"), pkg, ignore=shutil.ignore_patterns("*.so", "csrc", "__pycache__", "sdf2mesh"))
    code = ("import sdf2mesh_b200 as s2m\n"
            "try:\n    s2m.Sdf3DShader.from_source('fn sdf3d(p: vec3f) -> f32 { return 0.0; }')\n"
            "except ImportError as e:\n    print('IMPORTERROR', e)\n")
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=tmp_path, env={**os.environ, "PYTHONPATH": str(tmp_path)})
    assert "IMPORTERROR" in p.stdout and "no CPU fallback" in p.stdout, p.stdout + p.stderr
