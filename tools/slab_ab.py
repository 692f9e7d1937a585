#!/usr/bin/env python3
"""One z-slab of a grid on one GPU (what one rank of an N-GPU run does), under different chunking knobs.

usage: python tools/slab_ab.py mandelmesh2048 z_begin z_end [ENV=VAL,ENV=VAL ...]
Prints one JSON line per knob set: chunks, device time, host wall time (best of 6 after 3 warm-up runs).
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import sdf2mesh_b200 as s2m  # noqa: E402
from k1_ab import W, shader  # noqa: E402


def main():
    wl, zb, ze = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    f, res, bounds = W[wl]
    ctx = s2m.Context(0)
    mod = shader(f).create_shader_module(ctx)
    for spec in sys.argv[4:] or [""]:
        knobs = dict(kv.split("=") for kv in spec.split(",")) if spec else {}
        for k in list(os.environ):
            if k.startswith("S2M_") and k != "S2M_CACHE_DIR":
                os.environ.pop(k)
        os.environ.update(knobs)
        p, _ = s2m.params_from_cli(res, bounds, flags=s2m.MESH_QUADS_U32 | s2m.MESH_RELATIVE_QUADS)
        p.z_begin, p.z_end = zb, ze
        best = None
        for it in range(9):
            t0 = time.perf_counter()
            r = s2m.mesh_run(ctx, mod, p)
            wall = (time.perf_counter() - t0) * 1e3
            d = r.data()
            t = dict(d.timings)
            r.free()
            if it >= 3 and (best is None or wall < best[0]):
                best = (wall, t)
        print(json.dumps({"workload": wl, "z": [zb, ze], "knobs": knobs, "chunks": best[1].get("chunks"), "device_ms": round(best[1]["device_ms"], 3),
                          "wall_ms": round(best[0], 3)}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
