"""Multi-GPU z-slab decomposition: one process per GPU, no data-path collective.

Rank g meshes true cell slices [z_g, z_{g+1}) and recomputes the single slice z_g - 1 locally (its
one-plane corner halo) so it can name that slice's vertices in its quads.  The only exchange is
one all-gather of the per-slab vertex counts (8 bytes per rank, NCCL over NVLink when the backend
is nccl) whose exclusive prefix is the slab's global vertex base.  The
reference has no multi-device path at all (one adapter, one queue: main.rs:180-196).
"""
from typing import List, Optional, Sequence

import numpy as np


def n_scanned_slices(res_z: int, all_slices: bool) -> int:
    """slices the reference's loop actually reads back: res_z - 1 unless ALL_SLICES"""
    return res_z if all_slices else res_z - 1


def partition_slices(n_slices: int, world: int, cost: Optional[Sequence[float]] = None) -> List[int]:
    """Boundaries b[0..world] with b[0] = 0, b[world] = n_slices, strictly increasing.

    cost: relative cost of equal-thickness z bands (any length >= 1, e.g. from s2m_cost_probe);
    None = equal thickness.  Slab g gets ~1/world of the total cost (for the mandelbulb
    the work is concentrated in r <= 2, so equal thickness starves the outer ranks).
    """
    if world < 1:
        raise ValueError("world must be >= 1")
    if n_slices < world:
        # an empty slab cannot be expressed: z_begin == z_end == 0 means "the whole grid" to s2m_mesh_begin
        raise ValueError(f"{n_slices} z-slices cannot be split over {world} ranks (every rank needs at least one)")
    if cost is None or len(cost) == 0 or not np.isfinite(np.sum(cost)) or np.sum(cost) <= 0:
        return [int(round(n_slices * g / world)) for g in range(world + 1)]
    c = np.asarray(cost, np.float64)
    c = np.maximum(c, c.max() * 1e-3)  # every band costs something (launch + memory traffic)
    # piecewise-linear cumulative cost over slice index
    edges = np.linspace(0.0, float(n_slices), len(c) + 1)
    cum = np.concatenate([[0.0], np.cumsum(c)])
    targets = cum[-1] * np.arange(world + 1) / world
    b = np.interp(targets, cum, edges)
    out = [int(round(x)) for x in b]
    out[0], out[-1] = 0, n_slices
    for i in range(1, world + 1):   # strictly increasing: every slab has at least one slice, however skewed the profile
        out[i] = min(max(out[i], out[i - 1] + 1), n_slices - (world - i))
    return out


def exclusive_bases(counts: Sequence[int]) -> List[int]:
    out, acc = [], 0
    for c in counts:
        out.append(acc)
        acc += int(c)
    return out


_GATHER_BUF = {}


def allgather_counts(local_count: int, device=None, ctx=None) -> List[int]:
    """one all-gather of a single int64 per rank over torch.distributed (nccl or gloo).

    One collective into one tensor and one device->host read (a per-rank `.item()` costs a device
    synchronisation each: 8 of them took 2 ms of a 12 ms step on 8 GPUs).  With `ctx` (the engine
    context of this rank) the result is read through `Context.read_device_words` -- a one-warp kernel
    writing to mapped pinned memory on torch's current stream -- because a `.tolist()` is a D2H
    memcpy that queues on the copy engine behind the slab's vertex copy (another 2 ms)."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return [int(local_count)]
    world = dist.get_world_size()
    dev = device if device is not None else "cpu"
    key = (str(dev), world)
    if key not in _GATHER_BUF:
        _GATHER_BUF[key] = (torch.zeros(1, dtype=torch.int64, device=dev), torch.zeros(world, dtype=torch.int64, device=dev))
    src, dst = _GATHER_BUF[key]
    src.fill_(int(local_count))
    dist.all_gather_into_tensor(dst, src)
    if ctx is not None and dst.is_cuda and world <= 32:
        return ctx.read_device_words(dst.data_ptr(), world, torch.cuda.current_stream(dst.device).cuda_stream)
    return [int(x) for x in dst.tolist()]


def _band_density(n_slices: int, cost: Optional[Sequence[float]]) -> np.ndarray:
    """relative cost of every slice from a coarse band profile (uniform if there is none)"""
    if cost is None or len(cost) == 0 or not np.isfinite(np.sum(cost)) or np.sum(cost) <= 0:
        return np.ones(n_slices)
    c = np.asarray(cost, np.float64)
    c = np.maximum(c, c.max() * 1e-3)
    centres = (np.arange(len(c)) + 0.5) * n_slices / len(c)
    return np.interp(np.arange(n_slices) + 0.5, centres, c)


def rebalance(bounds: Sequence[int], seconds: Sequence[float], cost: Optional[Sequence[float]] = None,
              damping: float = 1.0) -> List[int]:
    """Refine z-slab boundaries from the time every rank actually took for its slab.

    Model: the cost of slice z is shape(z) * corr_g for z in rank g's current slab, where shape is
    the coarse probe profile (`cost`, uniform if None) and corr_g = seconds[g] / sum(shape over the
    slab) absorbs what the probe cannot see (vertex and quad work, copies).  The new boundaries cut
    the cumulative model cost into equal parts.  Meant for repeated meshing of the same SDF
    (animation frames, bench warm-up): measure one steady-state step, refine once.
    """
    world = len(bounds) - 1
    n = int(bounds[-1])
    sec = np.asarray(seconds, np.float64)
    widths = np.diff(np.asarray(bounds))
    if world < 2 or not np.all(np.isfinite(sec)) or sec.sum() <= 0 or np.any(widths <= 0) or np.any(sec <= 0):
        return [int(b) for b in bounds]
    dens = _band_density(n, cost)
    for g in range(world):
        sl = slice(int(bounds[g]), int(bounds[g + 1]))
        dens[sl] *= sec[g] / dens[sl].sum()
    cum = np.concatenate([[0.0], np.cumsum(dens)])
    targets = cum[-1] * np.arange(world + 1) / world
    nb = np.interp(targets, cum, np.arange(n + 1, dtype=np.float64))
    nb = np.asarray(bounds, np.float64) + damping * (nb - np.asarray(bounds, np.float64))
    out = [int(round(x)) for x in nb]
    out[0], out[-1] = 0, n
    for i in range(1, world + 1):
        out[i] = min(max(out[i], out[i - 1] + 1), n - (world - i))
    return out


def broadcast_floats(values: Optional[Sequence[float]], count: int, device=None) -> List[float]:
    """rank 0's `values` (length `count`) on every rank"""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return list(values)
    t = torch.zeros(count, dtype=torch.float64, device=device if device is not None else "cpu")
    if dist.get_rank() == 0:
        t[:] = torch.tensor(list(values), dtype=torch.float64)
    dist.broadcast(t, 0)
    return t.tolist()


def broadcast_boundaries(bounds: Optional[List[int]], world: int, device=None) -> List[int]:
    """rank 0 decides the partition (its cost probe); everyone uses the same one"""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return list(bounds)
    t = torch.zeros(world + 1, dtype=torch.int64, device=device if device is not None else "cpu")
    if dist.get_rank() == 0:
        t[:] = torch.tensor(bounds, dtype=torch.int64)
    dist.broadcast(t, 0)
    return [int(x) for x in t.tolist()]


def mesh_slab(ctx, module, params, z_begin: int, z_end: int, gather=allgather_counts, rank: int = 0, device=None):
    """Mesh one z-slab and return (MeshResult, global_vertex_base, all_counts).

    begin() runs K1..K4a and yields the local vertex count; the all-gather turns counts into the
    global base; finish(base) emits quads with global 64-bit indices.
    """
    from . import engine
    params.z_begin, params.z_end = int(z_begin), int(z_end)
    res = engine.mesh_begin(ctx, module, params)
    counts = gather(res.info().n_vertices, device, ctx) if gather is allgather_counts else gather(res.info().n_vertices)
    base = exclusive_bases(counts)[rank]
    res.finish(base)
    return res, base, counts


def write_mesh_gathered(positions, normals, quads, global_vertex_base: int, path, binary_stl: bool = False,
                        group=None, dst: int = 0) -> None:
    """One STL / PLY file from a one-process-per-GPU run: every rank hands in its z-slab as it sits in
    pinned host memory after finish(base) -- own vertices, quads with global indices -- and rank `dst`
    writes the file the single-GPU run would have written (engine.write_mesh_arrays; slabs are
    consecutive in rank order, so no halo copies are needed).

    The arrays travel host to host, so `group` must be a gloo group when the default backend is nccl
    (`torch.distributed.new_group(backend="gloo")`); sizes first, then three point-to-point messages
    per rank.  The reference has one device and one writer (mesh.rs:182); this is its multi-process form."""
    import torch
    import torch.distributed as dist
    from . import engine
    pos = np.ascontiguousarray(positions, np.float32).reshape(-1, 3)
    nrm = np.ascontiguousarray(normals if normals is not None else np.zeros_like(pos), np.float32).reshape(-1, 3)
    q = np.ascontiguousarray(quads, np.uint64).reshape(-1, 4)
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        engine.write_mesh_arrays([(pos, nrm, q, int(global_vertex_base))], path, binary_stl)
        return
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    mine = torch.tensor([pos.shape[0], q.shape[0], int(global_vertex_base)], dtype=torch.int64)
    sizes = [torch.zeros(3, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, mine, group=group)
    dst_global = dist.get_global_rank(group, dst) if group is not None else dst
    if rank != dst:
        for a in (pos, nrm, q.view(np.int64)):
            if a.size:
                dist.send(torch.from_numpy(a.reshape(-1)), dst_global, group=group)
    else:
        parts = []
        for r in range(world):
            nv, nq, base = (int(x) for x in sizes[r].tolist())
            if r == rank:
                parts.append((pos, nrm, q, base))
                continue
            src = dist.get_global_rank(group, r) if group is not None else r
            bufs = [torch.empty(nv * 3, dtype=torch.float32), torch.empty(nv * 3, dtype=torch.float32), torch.empty(nq * 4, dtype=torch.int64)]
            for b in bufs:
                if b.numel():
                    dist.recv(b, src, group=group)
            parts.append((bufs[0].numpy().reshape(-1, 3), bufs[1].numpy().reshape(-1, 3), bufs[2].numpy().view(np.uint64).reshape(-1, 4), base))
        engine.write_mesh_arrays(parts, path, binary_stl)
    dist.barrier(group=group)
