"""Digests of z-slab extracts of a mesh: the known-answer form for grids too large to run the CPU oracle on in full
(BASELINE config 5, 4096^3).  A slab [z0, z1) of true cell slices is described by its vertices (label keys, nibbles,
positions, normals, in key order) and by its quads AS CELL-KEY 4-TUPLES, restricted to quads whose four vertices all
lie inside the slab (a quad of the slab's first slice that names slice z0-1 is left out on both sides), sorted."""
import hashlib

import numpy as np

from tests.support.digest import canon_f32
from tests.support.tolerance import index_quads_to_keys


def stratified_pairs(n_slices, k):
    """k slab starts z (slab = [z, z+2)), evenly spread over the scanned slices 0 .. n_slices-1"""
    return sorted({min(n_slices - 2, max(1, int((i + 0.5) * n_slices / k))) for i in range(k)})


def slab_digest(keys, nibbles, positions, normals, quads, index_base=0):
    keys = np.ascontiguousarray(keys, np.uint64)
    kq = index_quads_to_keys(quads, keys, index_base)
    h = hashlib.sha256()
    for a in (keys, np.ascontiguousarray(nibbles, np.uint8), canon_f32(positions), canon_f32(normals), np.ascontiguousarray(kq, np.uint64)):
        h.update(a.tobytes())
    return {"n_vertices": int(len(keys)), "n_quads": int(len(kq)), "sha256": h.hexdigest()}


def extract_slab(keys, label_add, z0, z1):
    """index range [i0, i1) of the vertices of true slices [z0, z1) inside a whole-grid key array (ascending)"""
    lo = np.uint64((z0 + label_add) << 32)
    hi = np.uint64((z1 + label_add) << 32)
    return int(np.searchsorted(keys, lo)), int(np.searchsorted(keys, hi))
