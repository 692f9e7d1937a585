#!/usr/bin/env python3
"""A/B of K1 variants on one GPU, no torch: S2M_K1_PACKED = 0 (one corner per evaluation),
1 (f32x2 pairs), 2 (f32x2 pairs + packed sqrt refinement) for a few workloads.

Prints one JSON line per (workload, variant): K1 time with the GPU to itself (S2M_NO_CHUNK_OVERLAP=1),
the step's device time and host wall time with the normal chunk pipeline, and the mesh counts
(which must not depend on the variant).

usage: python tools/k1_ab.py [workload:variants[:ENV=VAL,ENV=VAL] ...]
       e.g.  mandelmesh2048:0,1,2 torus2048:0,2 mandelmesh2048:1:S2M_K1_MINBLOCKS=5,S2M_K1_ROWS=1
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sdf2mesh_b200 as s2m  # noqa: E402

W = {"mandelmesh2048": ("mandelmesh.frag", 2048, 5.0), "mandelmesh1024": ("mandelmesh.frag", 1024, 5.0), "torus2048": ("torus.sdf3d", 2048, 2.0),
     "martin_cube1024": ("martin_cube.sdf3d", 1024, 2.0), "martin_cube512": ("martin_cube.sdf3d", 512, 2.0), "p_key1024": ("p_key.sdf3d", 1024, 20.0),
     "p_key1024_b2": ("p_key.sdf3d", 1024, 2.0)}


def shader(f):
    p = os.path.join(ROOT, "examples", f)
    return s2m.Sdf3DShader.from_glsl_fragment_shader(p, "sdf") if f.endswith(".frag") else s2m.Sdf3DShader.from_path(p)


def run(ctx, mod, params, n):
    best = None
    for _ in range(n):
        t0 = time.perf_counter()
        r = s2m.mesh_run(ctx, mod, params)
        wall = (time.perf_counter() - t0) * 1e3
        d = r.data()
        t = dict(d.timings)
        t["wall_ms"] = wall
        t["counts"] = [int(len(d.keys)), int(len(d.quads)), int(d.n_invalid_quads)]
        r.free()
        if best is None or t["wall_ms"] < best["wall_ms"]:
            best = t
    return best


def main():
    specs = sys.argv[1:] or ["mandelmesh2048:0,1,2", "torus2048:0,1,2", "martin_cube1024:0,2", "p_key1024:0,2"]
    ctx = s2m.Context(0)
    for spec in specs:
        wl, variants, *extra = spec.split(":")
        knobs = dict(kv.split("=") for kv in extra[0].split(",")) if extra else {}
        for k in list(os.environ):   # every knob of the previous spec goes, whatever it was
            if k.startswith("S2M_") and k not in ("S2M_CACHE_DIR",):
                os.environ.pop(k, None)
        os.environ.update(knobs)
        f, res, bounds = W[wl]
        params, _ = s2m.params_from_cli(res, bounds, flags=s2m.MESH_QUADS_U32 | s2m.MESH_TIMINGS)
        for v in variants.split(","):
            if v == "d":  # the engine's default policy
                os.environ.pop("S2M_K1_PACKED", None)
            else:
                os.environ["S2M_K1_PACKED"] = v
            t0 = time.perf_counter()
            mod = shader(f).create_shader_module(ctx)
            jit = (time.perf_counter() - t0) * 1e3
            run(ctx, mod, params, 2)  # warm-up: buffers, pinned pool
            os.environ["S2M_NO_CHUNK_OVERLAP"] = "1"
            alone = run(ctx, mod, params, 2)
            del os.environ["S2M_NO_CHUNK_OVERLAP"]
            piped = run(ctx, mod, params, 4)
            print(json.dumps({"workload": wl, "S2M_K1_PACKED": v, "knobs": knobs, "packed": mod.packed, "jit_ms": round(jit, 1),
                              "k1_alone_ms": round(alone["k1_slab_ms"], 3), "k4a_alone_ms": round(alone["k4_vertices_ms"], 3),
                              "k2_k3_k4b_alone_ms": [round(alone.get(k, 0.0), 3) for k in ("k2_classify_ms", "k3_compact_ms", "k4_quads_ms")],
                              "device_alone_ms": round(alone["device_ms"], 3), "wall_alone_ms": round(alone["wall_ms"], 3), "chunks": piped.get("chunks"), "chunks_alone": alone.get("chunks"),
                              "device_ms": round(piped["device_ms"], 3), "wall_ms": round(piped["wall_ms"], 3),
                              "Gvoxel_per_s": round(res ** 3 / piped["wall_ms"] / 1e6, 1), "counts": piped["counts"]}), flush=True)
            del mod
    ctx.close()


if __name__ == "__main__":
    main()
