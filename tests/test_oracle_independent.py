"""A SECOND, independent restatement of the reference path, in vectorised numpy, against the C++ oracle.

The reference ships no golden vectors for this path and cannot be run here, so
the oracle under oracle/ is "parity unpinned".  What can be done without the reference binary is to
restate the same source a second time, separately, and require the two restatements to agree bit for
bit.  This file was written from the reference sources directly, not from oracle.cpp:

  /root/reference/src/bin/sdf2mesh/dualcontour.wgsl   cell_bounds :22-27, cell_new :29-43, sign nibble :57-69,
                                                      _cell_adapt / _cell_change :72-83, crossings :86-131, main :161-180
  /root/reference/src/sdf3d_normal.wgsl:4-10          tetrahedral gradient
  /root/reference/src/bin/sdf2mesh/main.rs:139-175    resolution / bounds / eps;  :298-356 slice loop, `p.3 > 0.0`,
                                                      queue.submit AFTER the read-back (one-slice lag)
  /root/reference/src/mesh.rs:224-226, :267-331       key, quads, binary search;  src/lib.rs:187-211 swap / validity
  /root/reference/src/lib.rs:115-118                  Bounds3D::cube

Every arithmetic step is one numpy float32 operation (IEEE, no contraction), in the order the WGSL
writes it.  SDFs: the torus example (+ - * / sqrt only), and two SDFs defined here and handed to the
oracle through its plug-in hook, so that no SDF transcription is shared either.
"""
import ctypes

import numpy as np
import pytest

import oracle
from tests.support.digest import f32_equal

F = np.float32


def length2(x, y):
    return np.sqrt(x * x + y * y)


def length3(x, y, z):
    return np.sqrt(x * x + y * y + z * z)


def sdf_torus(x, y, z):  # examples/torus.sdf3d + sdf3d_primitives.wgsl:32-36
    qx = length2(x, z) - F(0.5)
    return length2(qx, y) - F(0.2)


def sdf_two_spheres(x, y, z):  # min of two spheres, one off-centre: sign changes on all 12 edge kinds
    a = length3(x - F(0.3), y, z + F(0.1)) - F(0.45)
    b = length3(x + F(0.35), y - F(0.2), z) - F(0.3)
    return np.minimum(a, b)


def sdf_slab_with_hole(x, y, z):  # a box-like slab (max of planes) minus a cylinder: flat faces lie ON grid planes
    d = np.maximum(np.maximum(np.abs(x) - F(0.5), np.abs(y) - F(0.25)), np.abs(z) - F(0.5))
    hole = F(0.2) - length2(x, z)
    return np.maximum(d, hole)


def restate(sdf, res, bounds):
    """-> keys, nibbles, positions, normals, quads (after swap, valid only), n_invalid"""
    with np.errstate(all="ignore"):
        half = F(bounds) * F(0.5)  # Bounds3D::cube: center -/+ splat(a) * 0.5
        bmin, bmax = F(0.0) - half, F(0.0) + half
        v = F(res - 1)
        size = (bmax - bmin) / v
        idx = np.arange(res, dtype=np.float32)
        cmin_axis = bmin + size * idx          # cell_bounds: bounds.min + size * f32(pos)
        cmax_axis = cmin_axis + size           # min + size
        eps = F(0.0001)
        keys, nibbles, positions, normals = [], [], [], []
        Y0, X0 = np.meshgrid(cmin_axis, cmin_axis, indexing="ij")   # [y, x]
        Y1, X1 = np.meshgrid(cmax_axis, cmax_axis, indexing="ij")
        yy, xx = np.meshgrid(np.arange(res), np.arange(res), indexing="ij")
        for z in range(res - 1):               # the slice read back in iteration z+1 is slice z (main.rs:321-355)
            Z0 = np.full_like(X0, cmin_axis[z])
            Z1 = np.full_like(X0, cmax_axis[z])
            c000, c100 = sdf(X0, Y0, Z0), sdf(X1, Y0, Z0)
            c010, c110 = sdf(X0, Y1, Z0), sdf(X1, Y1, Z0)
            c001, c101 = sdf(X0, Y0, Z1), sdf(X1, Y0, Z1)
            c011, c111 = sdf(X0, Y1, Z1), sdf(X1, Y1, Z1)
            zero, one = np.zeros_like(X0), np.ones_like(X0)

            def adapt(v0, v1):
                return (F(0.0) - v0) / (v1 - v0)

            def change(a, b, x, y, zc):
                m = (a > 0) != (b > 0)
                return np.where(m, x, zero), np.where(m, y, zero), np.where(m, zc, zero)

            changes = [
                change(c000, c001, zero, zero, adapt(c000, c001)), change(c010, c011, zero, one, adapt(c010, c011)),
                change(c100, c101, one, zero, adapt(c100, c101)), change(c110, c111, one, one, adapt(c110, c111)),
                change(c000, c010, zero, adapt(c000, c010), zero), change(c001, c011, zero, adapt(c001, c011), one),
                change(c100, c110, one, adapt(c100, c110), zero), change(c101, c111, one, adapt(c101, c111), one),
                change(c000, c100, adapt(c000, c100), zero, zero), change(c001, c101, adapt(c001, c101), zero, one),
                change(c010, c110, adapt(c010, c110), one, zero), change(c011, c111, adapt(c011, c111), one, one)]
            ax, ay, az, cnt = zero.copy(), zero.copy(), zero.copy(), zero.copy()
            for cx, cy, cz in changes:
                use = (cx > 0) | (cy > 0) | (cz > 0)
                ax = np.where(use, ax + cx, ax)
                ay = np.where(use, ay + cy, ay)
                az = np.where(use, az + cz, az)
                cnt = np.where(use, cnt + F(1.0), cnt)
            has = ~(cnt <= F(1.0))
            px = X0 + (X1 - X0) * ax / cnt     # c.bounds.min + bounds_size(c.bounds) * avg / change_count
            py = Y0 + (Y1 - Y0) * ay / cnt
            pz = Z0 + (Z1 - Z0) * az / cnt
            f1 = sdf(px + eps, py - eps, pz - eps)   # v1 = ( 1,-1,-1)
            f2 = sdf(px - eps, py - eps, pz + eps)   # v2 = (-1,-1, 1)
            f3 = sdf(px - eps, py + eps, pz - eps)   # v3 = (-1, 1,-1)
            f4 = sdf(px + eps, py + eps, pz + eps)   # v4 = ( 1, 1, 1)
            nx = ((f1 + (-f2)) + (-f3)) + f4
            ny = (((-f1) + (-f2)) + f3) + f4
            nz = (((-f1) + f2) + (-f3)) + f4
            ln = length3(nx, ny, nz)
            nx, ny, nz = nx / ln, ny / ln, nz / ln
            signs = (c100 > 0) * 1 + (c010 > 0) * 2 + (c001 > 0) * 4 + (c000 > 0) * 8
            sel = has                             # shader stores w = 1 for a vertex; host takes p.3 > 0.0
            label = z + 1
            keys.append(xx[sel].astype(np.uint64) | (yy[sel].astype(np.uint64) << np.uint64(16)) | (np.uint64(label) << np.uint64(32)))
            nibbles.append(signs[sel].astype(np.uint8))
            positions.append(np.stack([px[sel], py[sel], pz[sel]], 1))
            normals.append(np.stack([nx[sel], ny[sel], nz[sel]], 1))
        keys = np.concatenate(keys)
        nibbles = np.concatenate(nibbles)
        positions = np.concatenate(positions).astype(np.float32)
        normals = np.concatenate(normals).astype(np.float32)
    # mesh.rs:267-331
    index = {int(k): i for i, k in enumerate(keys)}
    MISSING = -1

    def vi(x, y, z):
        return index.get(x | (y << 16) | (z << 32), MISSING)

    quads, invalid = [], 0
    for k, s in zip(keys.tolist(), nibbles.tolist()):
        x, y, z = k & 0xFFFF, (k >> 16) & 0xFFFF, k >> 32
        c0, c1, c2, c3 = bool(s & 1), bool(s & 2), bool(s & 4), bool(s & 8)
        cand = []
        if c0 != c3 and y > 0 and z > 0:
            cand.append(((vi(x, y - 1, z - 1), vi(x, y, z - 1), vi(x, y, z), vi(x, y - 1, z)), c0))
        if c1 != c3 and x > 0 and z > 0:
            cand.append(((vi(x - 1, y, z - 1), vi(x, y, z - 1), vi(x, y, z), vi(x - 1, y, z)), not c1))
        if c2 != c3 and x > 0 and y > 0:
            cand.append(((vi(x - 1, y - 1, z), vi(x, y - 1, z), vi(x, y, z), vi(x - 1, y, z)), c2))
        for q, swap in cand:
            if swap:
                q = q[::-1]
            if MISSING in q:
                invalid += 1
            else:
                quads.append(q)
    return keys, nibbles, positions, normals, np.array(quads, np.uint64).reshape(-1, 4), invalid


def compare(o, mine, what):
    keys, nibbles, positions, normals, quads, invalid = mine
    assert np.array_equal(o.keys, keys), f"{what}: active cells / order"
    assert np.array_equal(o.nibbles, nibbles), f"{what}: sign nibbles"
    assert f32_equal(o.positions, positions).all(), f"{what}: positions"
    assert f32_equal(o.normals, normals).all(), f"{what}: normals"
    assert o.n_invalid_quads == invalid, f"{what}: invalid quads {o.n_invalid_quads} vs {invalid}"
    assert np.array_equal(o.quads, quads), f"{what}: quads"
    assert len(keys) > 100


@pytest.mark.parametrize("res,bounds", [(32, 2.0), (64, 2.0), (32, 1.5)])
def test_torus_two_restatements_agree(res, bounds):
    o = oracle.mesh_run("torus", res, bounds)
    try:
        compare(o, restate(sdf_torus, res, bounds), f"torus {res}^3")
    finally:
        o.free()


@pytest.mark.parametrize("name,sdf,res,bounds", [("two_spheres", sdf_two_spheres, 24, 2.0), ("slab_with_hole", sdf_slab_with_hole, 32, 2.0)])
def test_plugged_in_sdf_two_restatements_agree(name, sdf, res, bounds):
    """the same numpy function evaluates one point at a time for the oracle (plug-in hook) and whole
    slices for the restatement above; the flat faces of the slab lie exactly on grid planes for res 32,
    which exercises the `> 0.0` conventions (a corner value of exactly 0 is inside)"""
    def scalar(x, y, z):
        with np.errstate(all="ignore"):
            return float(sdf(np.array([x], F), np.array([y], F), np.array([z], F))[0])

    cb = ctypes.CFUNCTYPE(ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float)(scalar)
    oracle.set_plugin(cb)
    try:
        o = oracle.mesh_run("plugin", res, bounds)
        try:
            compare(o, restate(sdf, res, bounds), f"{name} {res}^3")
        finally:
            o.free()
    finally:
        oracle.set_plugin(None)
