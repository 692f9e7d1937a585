#!/usr/bin/env python3
"""Small runs of every kernel path, meant to be executed under compute-sanitizer:

    compute-sanitizer --tool memcheck  python tools/sanitize_run.py
    compute-sanitizer --tool racecheck python tools/sanitize_run.py
    compute-sanitizer --tool initcheck python tools/sanitize_run.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import sdf2mesh_b200 as s2m  # noqa: E402


def main():
    ctx = s2m.Context(0)
    ex = os.path.join(ROOT, "examples")
    torus = s2m.Sdf3DShader.from_path(os.path.join(ex, "torus.sdf3d")).create_shader_module(ctx)
    bulb = s2m.Sdf3DShader.from_glsl_fragment_shader(os.path.join(ex, "mandelmesh.frag"), "sdf").create_shader_module(ctx)
    runs = 0
    for mod, bounds in ((torus, 2.0), (bulb, 5.0)):
        for dims in ((40, 40, 40), (70, 45, 33)):
            for flags in (0, s2m.MESH_CLASSIFY_FROM_SLAB, s2m.MESH_EXACT_DENSE, s2m.MESH_ALL_SLICES | s2m.MESH_KEEP_CANDIDATES | s2m.MESH_KEEP_INVALID,
                          s2m.MESH_NO_SLAB | s2m.MESH_QUADS_U32, s2m.MESH_RELATIVE_QUADS):
                if flags == s2m.MESH_EXACT_DENSE and dims[0] > 40:
                    continue
                for budget in (0, (dims[0] + 32) * (dims[1] + 1) * 4 * 4):
                    h = bounds / 2
                    p = s2m.make_params(dims, [-h] * 3, [h] * 3, flags=flags, slab_budget_bytes=budget)
                    r = s2m.mesh_run(ctx, mod, p)
                    d = r.data()
                    assert len(d.keys) > 0
                    r.free()
                    runs += 1
        # z-slabs with halo
        p, _ = s2m.params_from_cli(32, bounds)
        base = 0
        for zb, ze in ((0, 11), (11, 20), (20, 31)):
            p.z_begin, p.z_end = zb, ze
            r = s2m.mesh_begin(ctx, mod, p)
            n = r.info().n_vertices
            r.finish(base)
            base += n
            r.free()
            runs += 1
        mod.eval_points(np.zeros((1000, 3), np.float32))
        s2m.cost_probe(ctx, mod, p, 8)
    # several slabs behind s2m_multi (sharing this device; counts through host memory)
    from sdf2mesh_b200 import _capi
    mc = s2m.MultiContext([0, 0, 0], _capi.MULTI_NO_NCCL)
    compiled = s2m.Sdf3DShader.from_glsl_fragment_shader(os.path.join(ex, "mandelmesh.frag"), "sdf").create_shader_module(None)
    p, _ = s2m.params_from_cli(64, 5.0)
    for _ in range(3):
        for r in mc.mesh_run(compiled, p):
            r.free()
        runs += 1
    mc.close()
    print("sanitize_run: %d runs ok" % runs)
    ctx.close()


if __name__ == "__main__":
    main()
