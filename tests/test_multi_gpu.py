"""Several z-slabs behind ONE C-ABI call (s2m_multi_*, csrc/multi.cpp): the assembled mesh must be the one-slab mesh.

The reference has one device and one slice loop (main.rs:180-196, :298-356); its VertexList order is the key order
(mesh.rs:224-226), so concatenating the slabs' vertices in z order and adding each slab's base to its quads must
reproduce the single run bit for bit.  On a one-GPU box the slabs share the device (S2M_MULTI_NO_NCCL allows an
ordinal to repeat); with two or more GPUs the counts travel through ncclAllGather.
"""
import json
import os

import numpy as np
import pytest

import oracle
import sdf2mesh_b200 as s2m
from sdf2mesh_b200 import _capi
from tests.conftest import ROOT, load_example_shader
from tests.support.digest import mesh_digests

pytestmark = pytest.mark.gpu
GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "digests.json")))


def assemble(parts):
    """z-ordered slab results -> (positions, normals, keys, nibbles, global quads, invalid count)"""
    ds = [p.data() for p in parts]
    base = 0
    for p, d in zip(parts, ds):
        assert p.info().global_vertex_base == base   # exclusive prefix of the vertex counts
        assert d.quad_index_add == base              # S2M_MESH_RELATIVE_QUADS: the base is reported, not added
        base += len(d.keys)
    cat = lambda xs, dt, shape: np.concatenate([np.asarray(x, dt).reshape(shape) for x in xs])
    return (cat([d.positions for d in ds], np.float32, (-1, 3)), cat([d.normals for d in ds], np.float32, (-1, 3)),
            cat([d.keys for d in ds], np.uint64, (-1,)), cat([d.nibbles for d in ds], np.uint8, (-1,)),
            cat([d.global_quads() for d in ds], np.uint64, (-1, 4)), sum(d.n_invalid_quads for d in ds))


@pytest.mark.parametrize("name,res,bounds,slabs,flags", [
    ("torus", 64, 2.0, 3, 0), ("mandelbulb", 128, 5.0, 4, 0), ("mandelbulb", 128, 5.0, 2, s2m.MESH_QUADS_U32),
    ("p_key", 64, 2.0, 5, s2m.MESH_ALL_SLICES)])
def test_slabs_sharing_one_device_equal_the_oracle(ctx, name, res, bounds, slabs, flags):
    mc = s2m.MultiContext([0] * slabs, _capi.MULTI_NO_NCCL)
    assert len(mc) == slabs and mc.nccl_version == 0
    compiled = load_example_shader(name).create_shader_module(None)
    p, _ = s2m.params_from_cli(res, bounds, flags=flags)
    o = oracle.mesh_run(name, res, bounds, flags=flags & s2m.MESH_ALL_SLICES)
    seen = set()
    for run in range(4):   # run 2 refines the boundaries from measured times: the mesh must not notice
        parts = mc.mesh_run(compiled, p)
        b = mc.partition()
        seen.add(tuple(b))
        assert b[0] == 0 and b[-1] == (res if flags & s2m.MESH_ALL_SLICES else res - 1) and all(y > x for x, y in zip(b, b[1:]))
        pos, nrm, keys, nib, quads, ninv = assemble(parts)
        assert np.array_equal(keys, o.keys) and np.array_equal(nib, o.nibbles)
        assert np.array_equal(quads, o.quads) and ninv == o.n_invalid_quads
        assert np.array_equal(pos.view(np.uint32), np.asarray(o.positions, np.float32).view(np.uint32))
        assert np.array_equal(nrm.view(np.uint32), np.asarray(o.normals, np.float32).view(np.uint32))
        t = mc.timings()
        assert len(t["begin_ms"]) == slabs and t["wall_ms"] > 0
        for r in parts:
            r.free()
    o.free()
    mc.close()


def test_multi_writes_the_same_file(ctx, tmp_path):
    """s2m_write_mesh_parts over the multi run's results (relative quads) == the one-slab file"""
    mc = s2m.MultiContext([0, 0, 0], _capi.MULTI_NO_NCCL)
    compiled = load_example_shader("mandelbulb").create_shader_module(None)
    p, _ = s2m.params_from_cli(64, 5.0)
    parts = mc.mesh_run(compiled, p)
    one = s2m.mesh_run(ctx, compiled.instantiate(ctx), p)
    for ext in ("ply", "stl"):
        a, b = tmp_path / f"one.{ext}", tmp_path / f"many.{ext}"
        one.write_mesh(a)
        s2m.write_mesh_parts(parts, b)
        assert a.read_bytes() == b.read_bytes()
    for r in parts:
        r.free()
    one.free()
    mc.close()


def test_multi_rejects_what_it_cannot_do(ctx):
    with pytest.raises(s2m.S2mError):
        s2m.MultiContext([0, 0])            # NCCL cannot put two ranks on one device
    mc = s2m.MultiContext([0], 0)           # one device: no communicator needed
    assert mc.nccl_version == 0
    compiled = load_example_shader("torus").create_shader_module(None)
    p, _ = s2m.params_from_cli(32, 2.0)
    p.z_begin, p.z_end = 3, 9
    with pytest.raises(s2m.S2mError):
        mc.mesh_run(compiled, p)            # the multi run owns the z split
    p.z_begin = p.z_end = 0
    parts = mc.mesh_run(compiled, p)
    o = oracle.mesh_run("torus", 32, 2.0)
    assert np.array_equal(parts[0].data().keys, o.keys) and np.array_equal(parts[0].data().global_quads(), o.quads)
    parts[0].free()
    o.free()
    mc.close()
    mc4 = s2m.MultiContext([0] * 4, _capi.MULTI_NO_NCCL)
    p2, _ = s2m.params_from_cli(2, 2.0)     # one scanned slice for four slabs
    with pytest.raises(s2m.S2mError):
        mc4.mesh_run(compiled, p2)
    mc4.close()


def test_multi_context_with_changing_modules(ctx):
    """the per-device module cache is keyed by s2m_module_uid, not by the address of the compiled module: a module that is
    freed and another one compiled in its place (often at the same address) must not run the old kernels"""
    mc = s2m.MultiContext([0, 0], _capi.MULTI_NO_NCCL)
    p, _ = s2m.params_from_cli(64, 4.0)
    seen = set()
    for name in ("torus", "p_key", "torus", "mandelbulb", "p_key", "martin_cube"):
        compiled = load_example_shader(name).create_shader_module(None)
        seen.add(compiled.uid)
        parts = mc.mesh_run(compiled, p)
        o = oracle.mesh_run(name, 64, 4.0)
        pos, nrm, keys, nib, quads, ninv = assemble(parts)
        assert np.array_equal(keys, o.keys) and np.array_equal(quads, o.quads), name
        assert np.array_equal(pos.view(np.uint32), np.asarray(o.positions, np.float32).view(np.uint32)), name
        for r in parts:
            r.free()
        o.free()
        compiled.close()
        del compiled
    assert len(seen) == 6
    mc.close()


@pytest.mark.parametrize("key", ["mandelbulb_r512_b5_f0", "torus_r512_b2_f0"])
def test_nccl_count_exchange_over_all_gpus(key):
    """>= 2 GPUs: counts through ncclAllGather on communicators from ncclCommInitAll; golden digest of the assembled mesh"""
    import torch
    n = min(torch.cuda.device_count(), 8)
    if n < 2:
        pytest.skip("needs 2 GPUs")
    name, res, bounds = {"mandelbulb_r512_b5_f0": ("mandelbulb", 512, 5.0), "torus_r512_b2_f0": ("torus", 512, 2.0)}[key]
    mc = s2m.MultiContext(list(range(n)))
    assert mc.nccl_version >= 22000
    compiled = load_example_shader(name).create_shader_module(None)
    p, _ = s2m.params_from_cli(res, bounds)
    for run in range(3):
        parts = mc.mesh_run(compiled, p)
        pos, nrm, keys, nib, quads, ninv = assemble(parts)
        assert mesh_digests(pos, nrm, keys, nib, quads, ninv) == GOLDEN[key]
        for r in parts:
            r.free()
    mc.close()
