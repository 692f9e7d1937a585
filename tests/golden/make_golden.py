#!/usr/bin/env python3
"""Generate the known-answer set under tests/golden/ from the CPU oracle.

The reference ships no golden vectors for the meshing path and cannot be run in this image
(SURVEY.md section 4 / 8c), so these digests come from oracle/ (the line-by-line restatement), NOT
from the reference binary: parity is "unpinned" in the sense of the task statement.

usage: python tests/golden/make_golden.py [--big]      (--big adds the 512^3 cases; minutes of CPU)
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from tests.support.digest import mesh_digests  # noqa: E402

CASES = [
    # (sdf, resolution, bounds, flags)
    ("torus", 32, 2.0, 0), ("torus", 64, 2.0, 0), ("torus", 128, 2.0, 0), ("torus", 64, 2.0, 1),
    ("martin_cube", 64, 2.0, 0), ("martin_cube", 128, 2.0, 0),
    ("p_key", 64, 2.0, 0), ("p_key", 128, 20.0, 0),
    ("mandelbulb", 64, 5.0, 0), ("mandelbulb", 128, 5.0, 0), ("mandelbulb", 256, 5.0, 0),
    ("naga_sphere", 64, 2.5, 0),
]
BIG = [("torus", 512, 2.0, 0), ("martin_cube", 512, 2.0, 0), ("mandelbulb", 512, 5.0, 0), ("p_key", 512, 20.0, 0)]
# headline sizes: minutes of CPU on 16 cores; run on the GPU box's host (`--huge`), output copied back
HUGE = [("mandelbulb", 1024, 5.0, 0), ("mandelbulb", 2048, 5.0, 0), ("p_key", 1024, 20.0, 0),
        ("p_key", 1024, 2.0, 0),   # BASELINE config 3 as literally stated: --resolution 1024, default --bounds 2 (main.rs:150)
        ("torus", 2048, 2.0, 0)]   # bench.py's torus2048 entry (315 s on 8 cores)


# BASELINE config 5 in full: 68.7 G cells, 8 SDF evaluations each (~35 minutes on 8 cores, 5 GB); `--giant`
GIANT = [("mandelbulb", 4096, 5.0, 0)]


def key(c):
    return f"{c[0]}_r{c[1]}_b{c[2]:g}_f{c[3]}"


def main():
    path = os.path.join(HERE, "digests.json")
    out = json.load(open(path)) if os.path.exists(path) else {}
    if "--out" in sys.argv:
        path = sys.argv[sys.argv.index("--out") + 1]
    cases = CASES + (BIG if "--big" in sys.argv else []) + (HUGE if "--huge" in sys.argv else []) + (GIANT if "--giant" in sys.argv else [])
    for c in cases:
        if key(c) in out and "--force" not in sys.argv:
            continue
        t = time.time()
        m = oracle.mesh_run(c[0], c[1], c[2], flags=c[3])
        out[key(c)] = mesh_digests(m.positions, m.normals, m.keys, m.nibbles, m.quads, m.n_invalid_quads)
        print(key(c), out[key(c)]["n_vertices"], out[key(c)]["n_quads"], out[key(c)]["n_invalid_quads"], f"{time.time() - t:.1f}s", flush=True)
        if c == CASES[0]:
            np.savez_compressed(os.path.join(HERE, "torus_r32_b2.npz"), positions=m.positions, normals=m.normals, keys=m.keys,
                                nibbles=m.nibbles, quads=m.quads, n_invalid=np.int64(m.n_invalid_quads))
        m.free()
        json.dump(out, open(path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
