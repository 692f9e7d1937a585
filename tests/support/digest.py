"""Order-sensitive digests of mesh arrays (NaN payloads canonicalised: x86 and sm_100a produce
different NaN bit patterns for the same invalid operation)."""
import hashlib

import numpy as np


def canon_f32(a: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, np.float32).copy()
    u = a.view(np.uint32)
    u[np.isnan(a)] = 0x7FC00000
    return a


def digest(a: np.ndarray) -> str:
    a = np.ascontiguousarray(a)
    if a.dtype == np.float32:
        a = canon_f32(a)
    return hashlib.sha256(a.tobytes()).hexdigest()


def mesh_digests(positions, normals, keys, nibbles, quads, n_invalid):
    return {
        "n_vertices": int(len(keys)), "n_quads": int(len(quads)), "n_invalid_quads": int(n_invalid),
        "keys": digest(np.asarray(keys, np.uint64)), "nibbles": digest(np.asarray(nibbles, np.uint8)),
        "quads": digest(np.asarray(quads, np.uint64)), "positions": digest(np.asarray(positions, np.float32)),
        "normals": digest(np.asarray(normals, np.float32)),
    }


def f32_equal(a, b) -> np.ndarray:
    """bitwise equality with NaN == NaN"""
    a = np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)
    return (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))
