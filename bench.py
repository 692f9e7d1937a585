#!/usr/bin/env python3
"""bench.py -- headline benchmark: Gvoxels/s, end-to-end SDF -> quads, mandelmesh.frag 2048^3 --bounds 5.

    python bench.py --gpus N --steps K --warmup W            (N>1: under torchrun, one rank per GPU)
    python bench.py --impl reference ...                      (CPU restatement of the reference path)

A step = one complete pass of the hot path (K1 slab, K2 classify, K3 compact, K4 vertices + quads,
results copied into pinned host memory) over the whole grid.  At N>1 the grid is split into
contiguous z-slabs (cost-balanced from a coarse probe), every rank meshes its slab, the per-slab
vertex counts are exchanged with one NCCL all-gather, and quads carry global 64-bit indices; total
work is fixed, so "scaling" is "strong".

JSON keys beyond the base contract:
  value     voxels/s from the CUDA-event span first launch -> last kernel done (results in HBM)
  e2e       voxels/s from the host clock around the C-ABI calls, device->pinned-host copies included
  roofline  the dominant kernel (K1 slab) against measured HBM bandwidth, as the contract asks, plus
            `fp32`: the same kernel against the FP32 pipe (what actually bounds it for this SDF)
  kernels   per-kernel device time and achieved algorithmic GB/s
The CPU oracle is only used in the cpu_baseline / --impl reference legs (as the thing timed there).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (example file, kind, oracle sdf, resolution, bounds)
    "mandelmesh2048": ("mandelmesh.frag", "glsl", "mandelbulb", 2048, 5.0),
    "mandelmesh4096": ("mandelmesh.frag", "glsl", "mandelbulb", 4096, 5.0),  # beyond the reference's 2048 limit (BASELINE config 5)
    "mandelmesh1024": ("mandelmesh.frag", "glsl", "mandelbulb", 1024, 5.0),
    "mandelmesh512": ("mandelmesh.frag", "glsl", "mandelbulb", 512, 5.0),
    "torus2048": ("torus.sdf3d", "sdf3d", "torus", 2048, 2.0),
    "torus128": ("torus.sdf3d", "sdf3d", "torus", 128, 2.0),
    "martin_cube512": ("martin_cube.sdf3d", "sdf3d", "martin_cube", 512, 2.0),
    "p_key1024": ("p_key.sdf3d", "sdf3d", "p_key", 1024, 20.0),
}
# source-level flops per SDF evaluation, for the FP32 view of K1
FLOPS_PER_EVAL = {"mandelbulb": (170.0, 45.0, 0.27), "torus": (10.0, 10.0, 1.0), "martin_cube": (500.0, 500.0, 1.0), "p_key": (150.0, 150.0, 1.0)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: [self.lines.append((time.perf_counter(), l)) for l in self.proc.stdout], daemon=True).start()
        except OSError:
            self.proc = None

    def stop(self, t_from=0.0, t_to=float("inf")):
        """summarise the samples taken inside [t_from, t_to] (perf_counter clock)"""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, l in self.lines:
            if ts < t_from or ts > t_to + 0.15:
                continue
            f = [x.strip() for x in l.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "samples": len(sm), "reasons": sorted(reasons)}


def stratified_slices(n_slices, k):
    return sorted({int((i + 0.5) * n_slices / k) for i in range(k)})


def cpu_reference_sample(oracle_sdf, res, bounds, n_sample_slices):
    """Time the CPU restatement of the reference path on a bounded, z-stratified sample of the same
    workload (every host thread).  Returns (voxels/s, cores, seconds, description)."""
    import oracle
    slices = stratified_slices(res - 1, n_sample_slices)
    cores = oracle.lib().oracle_num_threads()
    t0 = time.perf_counter()
    nv = 0
    for z in slices:
        m = oracle.mesh_run(oracle_sdf, res, bounds, z_begin=z, z_end=z + 1)
        nv += len(m.keys)
        m.free()
    dt = time.perf_counter() - t0
    vox = len(slices) * res * res
    return vox / dt, cores, dt, f"{len(slices)} of {res - 1} z-slices of the {res}^3 grid, evenly spaced ({vox} voxels, {nv} vertices), 8 SDF evaluations per cell"


def run_reference_arm(args, wl):
    f, kind, osdf, res, bounds = WORKLOADS[wl]
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    for _ in range(args.warmup):
        cpu_reference_sample(osdf, res, bounds, 1)
    vals, secs = [], 0.0
    cores, desc = 1, ""
    for _ in range(args.steps):
        v, cores, dt, desc = cpu_reference_sample(osdf, res, bounds, args.ref_slices)
        vals.append(v); secs += dt
    value = sum(vals) / len(vals) / 1e9
    line = {
        "impl": "reference", "metric": "Gvoxels/s end-to-end SDF->quads", "value": value, "unit": "Gvoxel/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * secs / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic (analytic SDF, no random inputs)",
        "config": {"workload": wl, "sdf": f, "resolution": res, "bounds": bounds, "mode": "faithful (slices 0..R-2)"},
        "cpu_baseline": {"value": value, "unit": "Gvoxel/s", "cores": cores, "kind": "port", "sample": desc + "; extrapolated linearly in slice count"},
        "e2e": {"value": value, "unit": "Gvoxel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference (Rust + wgpu on lavapipe) cannot be built or run in this image; this is oracle/ (line-by-line CPU restatement) on the host cores",
    }
    print(json.dumps(line))


# dram__bytes_read.sum + dram__bytes_write.sum of ONE s2m_k1_slab launch (a 229-plane z-chunk of the
# default 4 GiB slab budget) from the ncu --set full capture summarised in profiles/r01c_k1_2048_full.md.
# Algorithmic bytes of that launch: 229 * 2049^2 * 4 B = 3.846 GB of corner values (+ 0.244 GB of
# corner-class planes that save K2 from re-reading the slab) -> no wasted DRAM traffic.
K1_TRAFFIC = {"mandelmesh2048": {"bytes_per_launch": 4.1174e9, "algorithmic_bytes_per_launch": 3.8457e9, "launches_per_step": 9,
                                 "source": "profiles/r01g_ncu_k1_packed_vs_scalar.md"}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="mandelmesh2048", choices=sorted(WORKLOADS))
    ap.add_argument("--ref-slices", type=int, default=64, help="z-slices per step of the CPU reference sample")
    ap.add_argument("--cpu-baseline-slices", type=int, default=192)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-balance", action="store_true", help="equal-thickness z-slabs instead of cost-balanced")
    ap.add_argument("--no-rebalance", action="store_true", help="keep the cost-probe partition; do not refine it from warm-up step times")
    ap.add_argument("--flags", type=int, default=0, help="extra S2M_MESH_* flags")
    ap.add_argument("--u64-quads", action="store_true", help="64-bit quad indices (default: u32, the reference's own index type)")
    ap.add_argument("--slab-budget-gb", type=float, default=0.0, help="bytes of corner slab resident at once (0 = engine default, 4 GiB chunks)")
    args = ap.parse_args()
    wl = args.workload
    if args.impl == "reference":
        run_reference_arm(args, wl)
        return

    import numpy as np
    import torch
    import sdf2mesh_b200 as s2m
    from sdf2mesh_b200 import distributed as dist_util

    f, kind, osdf, res, bounds = WORKLOADS[wl]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)

    ctx = s2m.Context(local_rank)
    path = os.path.join(ROOT, "examples", f)
    t0 = time.perf_counter()
    shader = s2m.Sdf3DShader.from_glsl_fragment_shader(path, "sdf") if kind == "glsl" else s2m.Sdf3DShader.from_path(path)
    module = shader.create_shader_module(ctx)
    jit_ms = (time.perf_counter() - t0) * 1e3
    if not args.u64_quads:
        args.flags |= s2m.MESH_QUADS_U32
    params, _ = s2m.params_from_cli(res, bounds, flags=args.flags)
    params.slab_budget_bytes = int(args.slab_budget_gb * (1 << 30))
    n_slices = dist_util.n_scanned_slices(res, bool(args.flags & s2m.MESH_ALL_SLICES))

    # z-slab partition (rank 0 probes the per-band cost, everyone uses its answer)
    bounds_z = None
    cost = None
    if world > 1:
        if not args.no_balance:
            cost = dist_util.broadcast_floats(s2m.cost_probe(ctx, module, params, 128) if rank == 0 else None, 128, dev)
        if rank == 0:
            bounds_z = dist_util.partition_slices(n_slices, world, cost)
        bounds_z = dist_util.broadcast_boundaries(bounds_z, world, dev)
    else:
        bounds_z = [0, n_slices]
    zb, ze = bounds_z[rank], bounds_z[rank + 1]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    phase = [0.0, 0.0, 0.0]  # this rank's begin / all-gather wait / finish, ms

    def step():
        if world == 1:
            r = s2m.mesh_run(ctx, module, params)  # one slab: quads are emitted and copied chunk by chunk
            counts = [r.info().n_vertices]
        else:
            # same sequence as distributed.mesh_slab, with the three phases timed on this rank
            params.z_begin, params.z_end = zb, ze
            t0 = time.perf_counter()
            r = s2m.mesh_begin(ctx, module, params)
            t1 = time.perf_counter()
            counts = dist_util.allgather_counts(r.info().n_vertices, dev, ctx)
            t2 = time.perf_counter()
            r.finish(dist_util.exclusive_bases(counts)[rank])
            t3 = time.perf_counter()
            phase[0] += (t1 - t0) * 1e3; phase[1] += (t2 - t1) * 1e3; phase[2] += (t3 - t2) * 1e3
        i = r.info()
        out = (i.n_vertices, i.n_quads, i.n_invalid_quads, {k[0]: getattr(i.timings, k[0]) for k in s2m._capi.Timings._fields_}, sum(counts), i.n_candidates)
        r.free()
        return out

    # nvidia-smi is started BEFORE the warm-up: its start-up stalls the driver for tens of ms
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.5)
    for w in range(args.warmup):
        step()
        if world > 1 and not args.no_balance and not args.no_rebalance and w == 1 and args.warmup >= 3:
            # One refinement, from the second warm-up step (the first one pays for allocations, and
            # the step after the refinement pays for re-allocations): what every rank's
            # begin() took, spread inside its slab according to the probe profile.
            # The all-gather in the middle of a step waits for the slowest begin(), so begin() is what
            # has to be equal across ranks (finish() is a quad copy, short and proportional to the vertices).
            own = torch.tensor([phase[0]], dtype=torch.float64, device=dev)
            allt = torch.zeros(world, dtype=torch.float64, device=dev)
            torch.distributed.all_gather_into_tensor(allt, own)
            bounds_z = dist_util.rebalance(bounds_z, allt.tolist(), cost)
            zb, ze = bounds_z[rank], bounds_z[rank + 1]
        phase[:] = [0.0, 0.0, 0.0]
    barrier()
    phase[:] = [0.0, 0.0, 0.0]
    t_start = time.perf_counter()
    stats = []
    for _ in range(args.steps):
        stats.append(step())
    barrier()
    t_end = time.perf_counter()
    wall = t_end - t_start
    if wall < 0.35:
        time.sleep(0.35 - wall)  # let at least a couple of 100 ms samples land
    clocks = sampler.stop(t_start, t_end)

    # The timed steps run K1 of chunk c+1 beside K3/K4a/K4b of chunk c (two streams), so K1's event span
    # contains those kernels' share of the machine.  Two extra UNTIMED steps with the overlap switched
    # off give every kernel's duration when it has the GPU to itself: that is what the roofline and the
    # ncu launch list (serialised by construction) are compared on.
    serial = None
    if world == 1:
        os.environ["S2M_NO_CHUNK_OVERLAP"] = "1"
        try:
            step()
            serial = step()[3]
        finally:
            del os.environ["S2M_NO_CHUNK_OVERLAP"]

    dev_ms = sum(s[3]["device_ms"] for s in stats)
    tot_ms = sum(s[3]["total_ms"] for s in stats)
    launches = sum(s[3]["launches"] for s in stats)
    nv, nq, ninv = stats[-1][0], stats[-1][1], stats[-1][2]
    ncand = stats[-1][5]
    d2h_bytes = nv * 33 + nq * (32 if args.u64_quads else 16)
    per = {k: sum(s[3][k] for s in stats) / args.steps for k in ("k1_slab_ms", "k2_classify_ms", "k3_compact_ms", "k4_vertices_ms", "k4_quads_ms", "d2h_ms")}
    per_rank = None
    if world > 1:
        mine = {"rank": rank, "slices": [zb, ze], "wall_ms_per_step": 1000.0 * wall / args.steps, "device_ms": dev_ms / args.steps,
                "vertices": int(nv), "candidates": int(ncand), "begin_ms": round(phase[0] / args.steps, 3),
                "allgather_wait_ms": round(phase[1] / args.steps, 3), "finish_ms": round(phase[2] / args.steps, 3), **{k: round(v, 3) for k, v in per.items()}}
        per_rank = [None] * world
        torch.distributed.all_gather_object(per_rank, mine)
        t = torch.tensor([wall, dev_ms, tot_ms, float(nv), float(nq), float(ninv), float(launches), float(d2h_bytes), float(ncand)] + [per[k] for k in sorted(per)],
                         dtype=torch.float64, device=dev)
        mx = t.clone(); torch.distributed.all_reduce(mx, op=torch.distributed.ReduceOp.MAX)
        sm = t.clone(); torch.distributed.all_reduce(sm, op=torch.distributed.ReduceOp.SUM)
        wall, dev_ms, tot_ms = mx[0].item(), mx[1].item(), mx[2].item()
        nv, nq, ninv, launches, d2h_bytes = int(sm[3].item()), int(sm[4].item()), int(sm[5].item()), int(sm[6].item()), int(sm[7].item())
        ncand = int(sm[8].item())
        per = {k: mx[9 + i].item() for i, k in enumerate(sorted(per))}
    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return

    voxels = float(res) ** 3
    hbm_peak, peak_src, sm_max = measured_peaks()
    value = voxels * args.steps / (dev_ms * 1e-3) / 1e9
    e2e = voxels * args.steps / wall / 1e9
    # K1 algorithmic bytes: one f32 per grid corner it writes (4 B/voxel slab write)
    planes = (ze - max(zb - 1, 0)) + 1 if world > 1 else n_slices + 1
    k1_bytes = 4.0 * (res + 1) * (res + 1) * planes
    k1_s = (serial["k1_slab_ms"] if serial else per["k1_slab_ms"]) * 1e-3
    # K2 reads K1's corner-class planes (8 B per 32 corners) and writes 1 bit per cell
    k2_bytes = (res + 1.0) * (((res + 1 + 31) // 32) * 8.0) * planes + (res * res * (planes - 1)) / 8.0
    full, early, frac_in = FLOPS_PER_EVAL.get(osdf, (100.0, 100.0, 1.0))
    flops = (res + 1.0) ** 2 * planes * (frac_in * full + (1 - frac_in) * early)
    fp32_peak = 148 * 128 * 2 * sm_max * 1e6 / 1e12
    traffic_known = wl in K1_TRAFFIC and world == 1 and not args.slab_budget_gb
    line = {
        "metric": "Gvoxels/s end-to-end SDF->quads", "value": value, "unit": "Gvoxel/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * wall / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic (analytic SDF, no random inputs)",
        "config": {"workload": wl, "sdf": f, "resolution": res, "bounds": bounds, "mode": "faithful (slices 0..R-2)" if not (args.flags & 1) else "all slices",
                   "quad_index": "u64" if args.u64_quads else "u32 (the reference's Quad(u32, u32, u32, u32))",
                   "parallelism": f"z-slabs x{world}" + ("" if world == 1 else (" equal" if args.no_balance else (" cost-probe" if args.no_rebalance else " cost-probe + refined from warm-up step times"))), "z_boundaries": bounds_z,
                   "k1": "packed f32x2: two corners per evaluation (FFMA2/FMUL2)" if module.packed else "one corner per evaluation",
                   "l2": "no L2 flush needed: the corner slab alone is %.1f GB per step, far larger than the 126 MB L2" % (k1_bytes / 1e9)},
        "mesh": {"candidates": ncand, "vertices": nv, "quads": nq, "invalid_quads": ninv, "triangles": 2 * nq, "Mtriangles_per_s": 2 * nq * args.steps / wall / 1e6},
        "e2e": {"value": e2e, "unit": "Gvoxel/s", "h2d_bytes_per_step": 64, "d2h_bytes_per_step": d2h_bytes},
        "gpu_launches": launches,
        "per_rank": per_rank,
        "clocks": clocks,
        "jit_ms": jit_ms,
        "roofline": {"kernel": "s2m_k1_slab", "bound": "hbm", "achieved": k1_bytes / k1_s / 1e9 if k1_s > 0 else None, "peak": hbm_peak, "unit": "GB/s",
                     "frac": (k1_bytes / k1_s / 1e9 / hbm_peak) if k1_s > 0 else None, "traffic": K1_TRAFFIC[wl]["bytes_per_launch"] if traffic_known else None, "traffic_detail": K1_TRAFFIC[wl] if traffic_known else None, "peak_source": peak_src,
                     "note": "K1 is FP32/issue bound for this SDF, not HBM bound (ncu: issue slots 83 % busy, FMA pipe 56 %, ALU pipe 51 %, DRAM 8 %); see fp32. "
                             "achieved = 4 B per corner written (SURVEY 8d) / K1 time with the GPU to itself (kernels_serialized; "
                             "in the timed steps K1 shares the machine with the previous chunk's K3/K4a/K4b, see kernels)",
                     "fp32": {"achieved_tflops_source_level": flops / k1_s / 1e12 if k1_s > 0 else None, "peak_tflops_nominal": fp32_peak,
                              "frac": flops / k1_s / 1e12 / fp32_peak if k1_s > 0 else None}},
        "kernels": {"k1_slab": {"ms": per["k1_slab_ms"], "GBps": k1_bytes / k1_s / 1e9 if k1_s > 0 else None},
                    "k2_classify": {"ms": per["k2_classify_ms"], "GBps": k2_bytes / (per["k2_classify_ms"] * 1e-3) / 1e9 if per["k2_classify_ms"] > 0 else None},
                    "k3_compact": {"ms": per["k3_compact_ms"]}, "k4_vertices": {"ms": per["k4_vertices_ms"]}, "k4_quads": {"ms": per["k4_quads_ms"]},
                    "d2h": {"ms": per["d2h_ms"]}, "device_total_ms": dev_ms / args.steps, "event_total_ms": tot_ms / args.steps,
                    "note": "event spans on each kernel's own stream; K1 (producer stream) overlaps K3/K4a/K4b (consumer stream), so the spans add up to more than device_total_ms"},
        "kernels_serialized": None if not serial else {
            "k1_slab_ms": serial["k1_slab_ms"], "k2_classify_ms": serial["k2_classify_ms"], "k3_compact_ms": serial["k3_compact_ms"],
            "k4_vertices_ms": serial["k4_vertices_ms"], "k4_quads_ms": serial["k4_quads_ms"], "device_ms": serial["device_ms"],
            "k1_share": serial["k1_slab_ms"] / serial["device_ms"],
            "note": "one untimed step with S2M_NO_CHUNK_OVERLAP=1: each kernel alone on the GPU, comparable with the ncu launch list in profiles/"},
    }
    if not args.no_cpu_baseline and world == 1:
        v, cores, dt, desc = cpu_reference_sample(osdf, res, bounds, args.cpu_baseline_slices)
        line["cpu_baseline"] = {"value": v / 1e9, "unit": "Gvoxel/s", "cores": cores, "kind": "port", "seconds": dt,
                                "sample": desc + "; extrapolated linearly in slice count"}
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
