"""Accuracy of the pinned f32 math (sdf2mesh_b200/csrc/s2m_math.h) against libm in double precision.

The header's functions are compiled here for the host; the GPU test test_math_gpu.py checks that the
device produces the same bits."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from tests.conftest import ROOT

FN = {"sin": 0, "cos": 1, "tan": 2, "asin": 3, "acos": 4, "atan": 5, "exp": 6, "exp2": 7, "log": 8, "log2": 9,
      "sinh": 10, "cosh": 11, "tanh": 12, "sqrt": 13, "abs": 14, "floor": 15, "fract": 16, "sign": 17, "round": 18,
      "asinh": 19, "acosh": 20, "atanh": 21}
FN2 = {"atan2": 100, "pow": 101, "min": 102, "max": 103, "div": 104, "fmod": 105, "mod": 106, "step": 107}


@pytest.fixture(scope="module")
def mathlib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("math") / "libmath_host.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-mfma", "-fPIC", "-shared", "-I", os.path.join(ROOT, "sdf2mesh_b200", "csrc"),
                           os.path.join(ROOT, "tests", "support", "math_host.cpp"), "-o", so])
    return ctypes.CDLL(so)


def map1(L, fn, x):
    x = np.ascontiguousarray(x, np.float32)
    y = np.empty_like(x)
    L.s2m_host_map1(fn, x.ctypes.data_as(ctypes.c_void_p), y.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(x.size))
    return y


def map2(L, fn, a, b):
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    y = np.empty_like(a)
    L.s2m_host_map2(fn, a.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p), y.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(a.size))
    return y


def max_ulp(y, ref):
    ref = np.asarray(ref, np.float64)
    with np.errstate(all="ignore"):
        ulp = np.maximum(np.spacing(np.abs(ref.astype(np.float32))).astype(np.float64), 2.0 ** -149)
        e = np.abs(y.astype(np.float64) - ref) / ulp
    ok = np.isfinite(ref) & np.isfinite(y) & (np.abs(ref) < 3.4e38)
    return float(e[ok].max())


RNG = np.random.default_rng(1)
N = 400_000
CASES = [
    ("sin", RNG.uniform(-20, 20, N), np.sin, 2.0), ("sin", RNG.uniform(-1e5, 1e5, N), np.sin, 2.0),
    ("sin", 10 ** RNG.uniform(5, 38, N), np.sin, 2.0), ("cos", RNG.uniform(-20, 20, N), np.cos, 2.0),
    ("cos", -(10 ** RNG.uniform(5, 38, N)), np.cos, 2.0), ("tan", RNG.uniform(-10, 10, N), np.tan, 3.0),
    ("asin", RNG.uniform(-1, 1, N), np.arcsin, 2.0), ("acos", RNG.uniform(-1, 1, N), np.arccos, 2.0),
    ("atan", RNG.uniform(-30, 30, N), np.arctan, 2.0), ("atan", 10 ** RNG.uniform(-30, 30, N), np.arctan, 2.0),
    ("exp", RNG.uniform(-103, 88, N), np.exp, 1.5), ("exp2", RNG.uniform(-126, 127, N), np.exp2, 1.5),
    ("log", 10 ** RNG.uniform(-44, 38, N), np.log, 1.5), ("log", RNG.uniform(0.5, 2, N), np.log, 1.5),
    ("log2", 10 ** RNG.uniform(-37, 38, N), np.log2, 2.0), ("sinh", RNG.uniform(-89, 89, N), np.sinh, 3.0),
    ("cosh", RNG.uniform(-89, 89, N), np.cosh, 3.0), ("tanh", RNG.uniform(-10, 10, N), np.tanh, 2.0),
]
RNG2 = np.random.default_rng(2)   # its own stream: the cases above keep the samples they were tuned on
CASES += [
    ("asinh", RNG2.uniform(-20, 20, N), np.arcsinh, 3.0), ("asinh", 10 ** RNG2.uniform(-30, 38, N), np.arcsinh, 3.0),
    ("asinh", -(10 ** RNG2.uniform(-8, 1, N)), np.arcsinh, 3.0), ("acosh", 1 + 10 ** RNG2.uniform(-7, 1, N), np.arccosh, 3.0),
    ("acosh", 10 ** RNG2.uniform(0, 38, N), np.arccosh, 3.0), ("atanh", RNG2.uniform(-1, 1, N), np.arctanh, 3.0),
    ("atanh", 10 ** RNG2.uniform(-30, 0, N), np.arctanh, 3.0), ("atanh", 1 - 10 ** RNG2.uniform(-7, 0, N), np.arctanh, 3.0),
]


@pytest.mark.parametrize("name,x,ref,tol", CASES, ids=[f"{c[0]}-{i}" for i, c in enumerate(CASES)])
def test_unary_accuracy(mathlib, name, x, ref, tol):
    x = x.astype(np.float32)
    with np.errstate(all="ignore"):
        y = map1(mathlib, FN[name], x)
        r = ref(x.astype(np.float64))
    assert max_ulp(y, r) <= tol


def test_pow_atan2_accuracy(mathlib):
    a = RNG.uniform(0, 4, N).astype(np.float32)
    b = RNG.uniform(-10, 10, N).astype(np.float32)
    with np.errstate(all="ignore"):
        assert max_ulp(map2(mathlib, FN2["pow"], a, b), np.power(a.astype(np.float64), b.astype(np.float64))) <= 2.0
        a2 = (10 ** RNG.uniform(-20, 20, N)).astype(np.float32)
        b2 = RNG.uniform(-3, 3, N).astype(np.float32)
        assert max_ulp(map2(mathlib, FN2["pow"], a2, b2), np.power(a2.astype(np.float64), b2.astype(np.float64))) <= 2.0
        y = RNG.uniform(-5, 5, N).astype(np.float32)
        x = RNG.uniform(-5, 5, N).astype(np.float32)
        assert max_ulp(map2(mathlib, FN2["atan2"], y, x), np.arctan2(y.astype(np.float64), x.astype(np.float64))) <= 2.0


def test_special_values(mathlib):
    inf, nan = np.inf, np.nan
    sp = np.array([0.0, -0.0, inf, -inf, nan, 1.0, -1.0], np.float32)

    def same(y, expect):
        e = np.array(expect, np.float32)
        assert np.array_equal(np.isnan(y), np.isnan(e))
        m = ~np.isnan(e)
        assert np.array_equal(y[m], e[m]) and np.array_equal(np.signbit(y[m]), np.signbit(e[m]))

    y = map1(mathlib, FN["sin"], sp)[:5]
    assert y[0] == 0.0 and y[1] == 0.0 and np.isnan(y[2:]).all()  # the sign of sin(-0) is not pinned
    same(map1(mathlib, FN["cos"], sp)[:5], [1.0, 1.0, nan, nan, nan])
    same(map1(mathlib, FN["atan"], sp)[:5], [0.0, -0.0, np.float32(np.pi / 2), -np.float32(np.pi / 2), nan])
    same(map1(mathlib, FN["exp"], sp)[:5], [1.0, 1.0, inf, 0.0, nan])
    same(map1(mathlib, FN["log"], sp), [-inf, -inf, inf, nan, nan, 0.0, nan])
    same(map1(mathlib, FN["asin"], np.array([2.0, -2.0], np.float32)), [nan, nan])
    same(map1(mathlib, FN["asinh"], sp), [0.0, -0.0, inf, -inf, nan, np.float32(np.arcsinh(1.0)), -np.float32(np.arcsinh(1.0))])
    same(map1(mathlib, FN["acosh"], np.array([1.0, 0.5, -3.0, inf, nan, -inf], np.float32)), [0.0, nan, nan, inf, nan, nan])
    same(map1(mathlib, FN["atanh"], np.array([0.0, -0.0, 1.0, -1.0, 1.5, -1.5, nan, inf], np.float32)), [0.0, -0.0, inf, -inf, nan, nan, nan, nan])
    # pow: C99 special cases
    A, B = np.meshgrid(sp, sp)
    with np.errstate(all="ignore"):
        ref = np.power(A.astype(np.float64), B.astype(np.float64)).astype(np.float32)
    got = map2(mathlib, FN2["pow"], A.ravel(), B.ravel()).reshape(A.shape)
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    assert np.array_equal(got[~np.isnan(ref)], ref[~np.isnan(ref)])
    # min/max: NaN ignored, -0 < +0
    same(map2(mathlib, FN2["min"], np.array([nan, 1.0, 0.0, -0.0], np.float32), np.array([2.0, nan, -0.0, 0.0], np.float32)), [2.0, 1.0, -0.0, -0.0])
    same(map2(mathlib, FN2["max"], np.array([nan, 1.0, 0.0, -0.0], np.float32), np.array([2.0, nan, -0.0, 0.0], np.float32)), [2.0, 1.0, 0.0, 0.0])
