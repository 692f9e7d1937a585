#!/usr/bin/env python3
"""Known-answer digests of sampled z-slabs (tests/support/slabs.py) from the CPU oracle, for the grids the oracle
cannot be run on in full within a test: BASELINE config 5 (mandelmesh.frag at 4096^3).  Like digests.json these
come from oracle/ (the line-by-line restatement), not from the reference binary.

usage: python tests/golden/make_slab_golden.py            (~3 minutes on 8 cores)
"""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from tests.support.slabs import slab_digest, stratified_pairs  # noqa: E402

CASES = [("mandelbulb", 4096, 5.0, 32), ("mandelbulb", 2048, 5.0, 16)]


def main():
    path = os.path.join(HERE, "slab_digests.json")
    out = json.load(open(path)) if os.path.exists(path) else {}
    for name, res, bounds, k in CASES:
        key = f"{name}_r{res}_b{bounds:g}_f0"
        if key in out and "--force" not in sys.argv:
            continue
        slabs = {}
        for z in stratified_pairs(res - 1, k):
            t = time.time()
            m = oracle.mesh_run(name, res, bounds, z_begin=z, z_end=z + 2)
            slabs[str(z)] = slab_digest(m.keys, m.nibbles, m.positions, m.normals, m.quads)
            print(key, z, slabs[str(z)]["n_vertices"], slabs[str(z)]["n_quads"], f"{time.time() - t:.1f}s", flush=True)
            m.free()
        out[key] = {"slices_per_slab": 2, "slabs": slabs}
        json.dump(out, open(path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
