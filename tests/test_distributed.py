"""Host-side multi-GPU logic on CPU: partitioning, count exchange (gloo, world_size 2), offsets."""
import os
import socket

import numpy as np
import pytest

from sdf2mesh_b200 import distributed as D


def test_partition_equal():
    assert D.partition_slices(2047, 1) == [0, 2047]
    b = D.partition_slices(2047, 8)
    assert b[0] == 0 and b[-1] == 2047 and all(x <= y for x, y in zip(b, b[1:]))
    assert max(y - x for x, y in zip(b, b[1:])) - min(y - x for x, y in zip(b, b[1:])) <= 1


def test_partition_cost_weighted():
    cost = np.ones(128)
    cost[32:96] = 20.0  # all the work in the middle half
    b = D.partition_slices(1024, 4, cost)
    assert b[0] == 0 and b[-1] == 1024 and all(x <= y for x, y in zip(b, b[1:]))
    widths = [y - x for x, y in zip(b, b[1:])]
    assert widths[0] > widths[1] and widths[3] > widths[2], "outer slabs must be thicker than inner ones"
    edges = np.linspace(0, 1024, 129)
    cum = np.concatenate([[0], np.cumsum(cost)])
    per = np.diff(np.interp(b, edges, cum))
    assert per.max() / per.min() < 1.15
    assert D.partition_slices(10, 4, [0, 0, 0]) == D.partition_slices(10, 4)  # degenerate cost -> equal


def test_partition_never_yields_an_empty_slab():
    """z_begin == z_end == 0 means "whole grid" to s2m_mesh_begin, so an empty first slab must not come out of
    the partitioner however skewed the cost profile is"""
    cost = np.zeros(128)
    cost[-1] = 1.0  # everything in the last band
    b = D.partition_slices(64, 8, cost)
    assert b[0] == 0 and b[-1] == 64 and all(y > x for x, y in zip(b, b[1:])), b
    b = D.partition_slices(8, 8, cost)
    assert b == list(range(9))
    with pytest.raises(ValueError):
        D.partition_slices(7, 8)


def test_exclusive_bases_and_slice_count():
    assert D.exclusive_bases([5, 0, 7]) == [0, 5, 5]
    assert D.n_scanned_slices(2048, False) == 2047 and D.n_scanned_slices(2048, True) == 2048


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    counts = D.allgather_counts(1000 * (rank + 1) + 7)
    bounds = D.broadcast_boundaries([0, 300, 511] if rank == 0 else None, world)
    q.put((rank, counts, D.exclusive_bases(counts)[rank], bounds))
    dist.destroy_process_group()


def test_count_exchange_gloo_world2():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert out[0] == (0, [1007, 2007], 0, [0, 300, 511])
    assert out[1] == (1, [1007, 2007], 1007, [0, 300, 511])


def _writer_worker(rank, world, port, path, q):
    """each rank owns a z-slab of one oracle mesh, split the way the engine's slabs are (own vertices,
    quads of the cells it emits, global indices); rank 0 writes"""
    import torch.distributed as dist
    import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    o = oracle.mesh_run("torus", 32, 2.0)
    z = (o.keys >> np.uint64(32)).astype(np.int64)
    cuts = [0] + [int(np.quantile(z, (g + 1) / world)) for g in range(world - 1)] + [1 << 20]
    emit = np.maximum.reduce([z[o.quads[:, k].astype(np.int64)] for k in range(4)])
    own_v = (z >= cuts[rank]) & (z < cuts[rank + 1])
    own_q = (emit >= cuts[rank]) & (emit < cuts[rank + 1])
    base = int((z < cuts[rank]).sum())
    for ext, binary in (("stl", False), ("ply", False), ("bin.stl", True)):
        D.write_mesh_gathered(o.positions[own_v], o.normals[own_v], o.quads[own_q], base, f"{path}.{ext}", binary_stl=binary)
    q.put((rank, int(own_v.sum()), int(own_q.sum())))
    o.free()
    dist.destroy_process_group()


def test_gathered_writer_gloo_world2(tmp_path):
    """one-process-per-GPU output path on CPU: two ranks' z-slabs -> the file of the whole mesh"""
    import oracle
    import sdf2mesh_b200 as s2m
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_writer_worker, args=(r, 2, port, str(tmp_path / "g"), q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    o = oracle.mesh_run("torus", 32, 2.0)
    assert out[0][1] > 0 and out[1][1] > 0 and out[0][1] + out[1][1] == len(o.keys) and out[0][2] + out[1][2] == len(o.quads)
    o.write_stl(tmp_path / "ref.stl")
    o.write_ply(tmp_path / "ref.ply")
    assert (tmp_path / "g.stl").read_bytes() == (tmp_path / "ref.stl").read_bytes()
    assert (tmp_path / "g.ply").read_bytes() == (tmp_path / "ref.ply").read_bytes()
    s2m.write_mesh_arrays([(o.positions, None, o.quads)], tmp_path / "ref.bin.stl", binary_stl=True)
    assert (tmp_path / "g.bin.stl").read_bytes() == (tmp_path / "ref.bin.stl").read_bytes()
    o.free()


def test_rebalance_from_measured_times():
    b = [0, 406, 678, 864, 1023, 1183, 1368, 1640, 2047]
    t = [6.75, 8.3, 10.7, 10.3, 10.45, 10.1, 7.9, 6.8]
    nb = D.rebalance(b, t)
    assert nb[0] == 0 and nb[-1] == 2047 and all(x < y for x, y in zip(nb, nb[1:]))
    rate = np.array(t) / np.diff(b)
    pred = []
    for g in range(8):
        z = np.arange(nb[g], nb[g + 1])
        pred.append(rate[np.searchsorted(b, z, side="right") - 1].sum())
    assert max(pred) / min(pred) < 1.02 and max(pred) < max(t) * 0.9
    assert D.rebalance([0, 10], [1.0]) == [0, 10]
    assert D.rebalance(b, [0.0] * 8) == b
    # with a probe profile the cost inside a slab follows the profile; the fastest ranks (the outer
    # slabs) get more slices, the slowest (middle) fewer
    z = (np.arange(128) + 0.5) / 128 * 2 - 1
    profile = 1.0 + 20.0 * np.exp(-(z / 0.35) ** 2)
    b0 = D.partition_slices(2047, 8, profile)
    t2 = [6.75, 8.3, 10.7, 10.3, 10.45, 10.1, 7.9, 6.8]
    with_profile = D.rebalance(b0, t2, profile)
    uniform = D.rebalance(b0, t2)
    assert with_profile[1] > b0[1] and uniform[1] > b0[1]
    assert with_profile[4] - with_profile[3] < b0[4] - b0[3]
    assert all(x < y for x, y in zip(with_profile, with_profile[1:]))


def test_c_abi_partition_matches_the_python_one(built):
    """s2m_partition_slices / s2m_rebalance_slices (csrc/multi.cpp, used by s2m_multi_mesh_run and the CLI's --gpus N)
    follow distributed.partition_slices / rebalance: same boundaries up to a tie in the rounding, always strictly
    increasing, always covering [0, n]"""
    from sdf2mesh_b200 import multi as M
    rng = np.random.default_rng(0)
    for t in range(200):
        n, w = int(rng.integers(8, 3000)), int(rng.integers(1, 9))
        cost = None if t % 5 == 0 else rng.uniform(0, 1, int(rng.integers(1, 200))) ** 3
        a, b = D.partition_slices(n, w, cost), M.partition_slices(n, w, cost)
        assert b[0] == 0 and b[-1] == n and all(y > x for x, y in zip(b, b[1:])), b
        assert max(abs(x - y) for x, y in zip(a, b)) <= 1, (n, w, a, b)
        if w >= 2:
            sec = rng.uniform(0.5, 2, w)
            a2, b2 = D.rebalance(a, sec, cost), M.rebalance_slices(a, sec, cost)
            assert b2[0] == 0 and b2[-1] == n and all(y > x for x, y in zip(b2, b2[1:])), b2
            assert max(abs(x - y) for x, y in zip(a2, b2)) <= 1, (n, w, a, a2, b2)
    with pytest.raises(Exception):
        M.partition_slices(3, 8)
    # a slab that took twice as long gets thinner
    b = M.rebalance_slices([0, 100, 200], [2.0, 1.0])
    assert b[1] < 100


def test_multi_context_needs_a_device(built):
    """no CPU fallback behind s2m_multi_create either"""
    import sdf2mesh_b200 as s2m
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(s2m.S2mError) as e:
        s2m.MultiContext([0, 1])
    assert e.value.kind == "NO_DEVICE"
