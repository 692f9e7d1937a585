#!/usr/bin/env python3
"""Derive the polynomial coefficients used by sdf2mesh_b200/csrc/s2m_math.h.

Near-minimax fits (Lawson-reweighted least squares in float64 on Chebyshev nodes),
rounded to float32.  The header pins these numbers; this script only documents where
they come from.  Accuracy of the resulting float32 functions is measured separately by
tests/test_math.py against libm in double precision.
"""
import numpy as np

def cheb_nodes(a, b, n):
    k = np.arange(n)
    x = np.cos(np.pi * (k + 0.5) / n)
    return 0.5 * (a + b) + 0.5 * (b - a) * x

def lawson_fit(basis, target, weight, iters=60):
    """min max |weight * (basis @ c - target)| approximately."""
    w = np.ones_like(target)
    c = None
    for _ in range(iters):
        sw = np.sqrt(w) * weight
        c, *_ = np.linalg.lstsq(basis * sw[:, None], target * sw, rcond=None)
        err = np.abs(weight * (basis @ c - target))
        w = w * (err / err.max() + 1e-3)
        w /= w.sum()
    err = np.abs(weight * (basis @ c - target))
    return c, err.max()

def show(name, c, err):
    c32 = np.asarray(c, dtype=np.float32)
    print(f"// {name}: max weighted err {err:.3e}")
    for i, v in enumerate(c32):
        print(f"//   c{i} = {float(v):.9e}f  ({float(v).hex()})")
    print()

N = 4000

# sin(r) = r + r^3 * S(s), s = r^2, r in [-pi/4, pi/4]
r = cheb_nodes(1e-4, np.pi / 4 * 1.01, N)
s = r * r
tgt = (np.sin(r) - r) / r**3
B = np.stack([s**k for k in range(4)], axis=1)
c, e = lawson_fit(B, tgt, (r**3) / np.sin(r))
show("sin S(s) deg3", c, e)

# cos(r) = 1 - s/2 + s^2 * C(s)
tgt = (np.cos(r) - 1 + s / 2) / s**2
B = np.stack([s**k for k in range(4)], axis=1)
c, e = lawson_fit(B, tgt, s**2 / np.cos(r))
show("cos C(s) deg3", c, e)

# tan(r) = r + r^3 * T(s)
tgt = (np.tan(r) - r) / r**3
B = np.stack([s**k for k in range(7)], axis=1)
c, e = lawson_fit(B, tgt, r**3 / np.tan(r))
show("tan T(s) deg6", c, e)

# atan(t) = t + t^3 * A(s), t in [0, 1]
t = cheb_nodes(1e-4, 1.0, N)
s = t * t
tgt = (np.arctan(t) - t) / t**3
for deg in (7, 8):
    B = np.stack([s**k for k in range(deg + 1)], axis=1)
    c, e = lawson_fit(B, tgt, t**3 / np.arctan(t))
    show(f"atan A(s) deg{deg}", c, e)

# asin(x) = x + x^3 * P(s), |x| <= 0.5
x = cheb_nodes(1e-4, 0.5, N)
s = x * x
tgt = (np.arcsin(x) - x) / x**3
for deg in (4, 5):
    B = np.stack([s**k for k in range(deg + 1)], axis=1)
    c, e = lawson_fit(B, tgt, x**3 / np.arcsin(x))
    show(f"asin P(s) deg{deg}", c, e)

# log1p(f) = f - f^2/2 + f^3 * L(f), f in [-1/3, 1/3]   (m in [2/3, 4/3])
f = cheb_nodes(-1.0 / 3, 1.0 / 3, N)
f = f[np.abs(f) > 1e-3]
tgt = (np.log1p(f) - f + f * f / 2) / f**3
for deg in (6, 7):
    B = np.stack([f**k for k in range(deg + 1)], axis=1)
    c, e = lawson_fit(B, tgt, np.abs(f**3 / np.log1p(f)))
    show(f"log1p L(f) deg{deg}", c, e)

# exp(f) = 1 + f + f^2 * E(f), f in [-ln2/2, ln2/2]
f = cheb_nodes(-np.log(2) / 2 * 1.01, np.log(2) / 2 * 1.01, N)
f = f[np.abs(f) > 1e-3]
tgt = (np.exp(f) - 1 - f) / f**2
for deg in (4, 5):
    B = np.stack([f**k for k in range(deg + 1)], axis=1)
    c, e = lawson_fit(B, tgt, f**2 / np.exp(f))
    show(f"exp E(f) deg{deg}", c, e)

# exp2(f) = 1 + f * G(f), f in [-0.5, 0.5]
f = cheb_nodes(-0.5, 0.5, N)
f = f[np.abs(f) > 1e-3]
tgt = (np.exp2(f) - 1) / f
for deg in (5, 6):
    B = np.stack([f**k for k in range(deg + 1)], axis=1)
    c, e = lawson_fit(B, tgt, np.abs(f) / np.exp2(f))
    show(f"exp2 G(f) deg{deg}", c, e)

# atanh(q) = q + q^3 * H(s), s = q^2, q in [-(sqrt2-1)^2.., ] : m in [sqrt(.5), sqrt(2)] -> |q| <= 0.17158
q = cheb_nodes(1e-4, 0.1716, N)
s = q * q
tgt = (np.arctanh(q) - q) / q**3
for deg in (3, 4):
    B = np.stack([s**k for k in range(deg + 1)], axis=1)
    c, e = lawson_fit(B, tgt, q**3 / np.arctanh(q))
    show(f"atanh H(s) deg{deg}", c, e)

# tanh(a) = a + a^3 * P(s), a in [0, 0.55]
a = cheb_nodes(1e-4, 0.55, N)
s = a * a
tgt = (np.tanh(a) - a) / a**3
for deg in (5, 6):
    B = np.stack([s**k for k in range(deg + 1)], axis=1)
    c, e = lawson_fit(B, tgt, a**3 / np.tanh(a))
    show(f"tanh P(s) deg{deg}", c, e)

# sinh(a) = a + a^3 * P(s), a in [0, 1]
a = cheb_nodes(1e-4, 1.0, N)
s = a * a
tgt = (np.sinh(a) - a) / a**3
B = np.stack([s**k for k in range(4)], axis=1)
c, e = lawson_fit(B, tgt, a**3 / np.sinh(a))
show("sinh P(s) deg3", c, e)
