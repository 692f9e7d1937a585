"""CPU oracle bindings -- TEST INFRASTRUCTURE ONLY (see oracle/oracle.cpp).

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs; never from sdf2mesh_b200/.
"""
import ctypes
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

SDF_IDS = {"torus": 0, "martin_cube": 1, "p_key": 2, "mandelbulb": 3, "naga_sphere": 4, "plugin": 99}
FLAG_ALL_SLICES = 1
FLAG_CONSISTENT_CORNERS = 64  # not the reference's arithmetic; mirrors S2M_MESH_CONSISTENT_CORNERS


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("oracle.cpp", "sdf_examples.h")]
    srcs.append(os.path.join(_HERE, "..", "sdf2mesh_b200", "csrc", "s2m_math.h"))
    stale = force or not os.path.exists(so) or any(
        os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if stale:
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "liboracle.so"])
    return so


def build_indep(force: bool = False) -> str:
    so = os.path.join(_HERE, "libindep.so")
    src = os.path.join(_HERE, "indep.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(src) > os.path.getmtime(so):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libindep.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        L.oracle_mesh_run.restype = ctypes.c_void_p
        L.oracle_mesh_run.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                      ctypes.c_float, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int]
        L.oracle_mesh_counts.argtypes = [ctypes.c_void_p] + [ctypes.c_void_p] * 5
        L.oracle_mesh_copy.argtypes = [ctypes.c_void_p] * 6
        L.oracle_mesh_free.argtypes = [ctypes.c_void_p]
        L.oracle_mesh_invalid.restype = ctypes.c_uint64
        L.oracle_mesh_invalid.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.oracle_write_stl.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
        L.oracle_write_ply.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
        L.oracle_eval.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64]
        L.oracle_set_plugin.argtypes = [ctypes.c_void_p]
        L.oracle_axis_coords.argtypes = [ctypes.c_uint32, ctypes.c_float, ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p]
        L.oracle_rust_f32.restype = ctypes.c_char_p
        L.oracle_rust_f32.argtypes = [ctypes.c_float]
        _LIB = L
    return _LIB


@dataclass
class OracleMesh:
    positions: np.ndarray   # (N,3) f32
    normals: np.ndarray     # (N,3) f32
    keys: np.ndarray        # (N,) u64   x | y<<16 | label<<32   (mesh.rs:224-226)
    nibbles: np.ndarray     # (N,) u8    bit0 s100, bit1 s010, bit2 s001, bit3 s000
    quads: np.ndarray       # (Q,4) u64  after swap, emission order, valid only
    n_invalid_quads: int
    seconds_cells: float
    seconds_quads: float
    _handle: int = 0

    def free(self):
        if self._handle:
            lib().oracle_mesh_free(self._handle)
            self._handle = 0

    def invalid_records(self) -> np.ndarray:
        """(n, 6) u64: key, edge, q0..q3 (UINT64_MAX = missing), in the reference's warning order"""
        n = lib().oracle_mesh_invalid(self._handle, None)
        out = np.empty((n, 6), np.uint64)
        if n:
            lib().oracle_mesh_invalid(self._handle, out.ctypes.data)
        return out

    def write_stl(self, path):
        assert lib().oracle_write_stl(self._handle, str(path).encode()) == 0

    def write_ply(self, path):
        assert lib().oracle_write_ply(self._handle, str(path).encode()) == 0


def cube_bounds(b: float):
    """lib.rs:115-118 Bounds3D::cube(a, 0): v = a*0.5 (f32); min = 0 - v, max = 0 + v."""
    v = np.float32(b) * np.float32(0.5)
    return (np.zeros(3, np.float32) - v).astype(np.float32), (np.zeros(3, np.float32) + v).astype(np.float32)


def mesh_run(sdf, res, bounds=2.0, eps=1e-4, flags=0, z_begin=0, z_end=0, threads=0, bmin=None, bmax=None) -> OracleMesh:
    L = lib()
    sid = SDF_IDS[sdf] if isinstance(sdf, str) else int(sdf)
    res3 = np.array([res] * 3 if np.isscalar(res) else res, dtype=np.uint32)
    if bmin is None:
        bmin, bmax = cube_bounds(bounds)
    bmin = np.ascontiguousarray(bmin, np.float32)
    bmax = np.ascontiguousarray(bmax, np.float32)
    h = L.oracle_mesh_run(sid, res3.ctypes.data, bmin.ctypes.data, bmax.ctypes.data, ctypes.c_float(eps),
                          flags, z_begin, z_end, threads)
    nv, nq, ninv = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()
    sc, sq = ctypes.c_double(), ctypes.c_double()
    L.oracle_mesh_counts(h, ctypes.byref(nv), ctypes.byref(nq), ctypes.byref(ninv), ctypes.byref(sc), ctypes.byref(sq))
    pos = np.empty((nv.value, 3), np.float32)
    nrm = np.empty((nv.value, 3), np.float32)
    keys = np.empty(nv.value, np.uint64)
    nib = np.empty(nv.value, np.uint8)
    quads = np.empty((nq.value, 4), np.uint64)
    L.oracle_mesh_copy(h, pos.ctypes.data, nrm.ctypes.data, keys.ctypes.data, nib.ctypes.data, quads.ctypes.data)
    return OracleMesh(pos, nrm, keys, nib, quads, ninv.value, sc.value, sq.value, h)


def set_plugin(fn_ptr) -> None:
    """SDF id "plugin" = this host function `float f(float x, float y, float z)` (a ctypes function or
    an address): the front-end's emitted code compiled for the host, see tests/support/host_eval.py"""
    addr = ctypes.cast(fn_ptr, ctypes.c_void_p) if fn_ptr is not None else ctypes.c_void_p(0)
    lib().oracle_set_plugin(addr)


def eval_points(sdf, pts) -> np.ndarray:
    sid = SDF_IDS[sdf] if isinstance(sdf, str) else int(sdf)
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 3)
    out = np.empty(pts.shape[0], np.float32)
    lib().oracle_eval(sid, pts.ctypes.data, out.ctypes.data, pts.shape[0])
    return out


def axis_coords(res, bmin, bmax):
    a = np.empty(res, np.float32)
    b = np.empty(res, np.float32)
    lib().oracle_axis_coords(res, ctypes.c_float(bmin), ctypes.c_float(bmax), a.ctypes.data, b.ctypes.data)
    return a, b


def rust_f32(x) -> str:
    return lib().oracle_rust_f32(ctypes.c_float(x)).decode()


# ---------------------------------------------------------------- independent evaluation (indep.cpp)
_INDEP = None


def indep_lib():
    global _INDEP
    if _INDEP is None:
        L = ctypes.CDLL(build_indep())
        L.indep_run.restype = ctypes.c_void_p
        L.indep_run.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32,
                                ctypes.c_uint32, ctypes.c_double, ctypes.c_int, ctypes.c_int]
        L.indep_count.restype = ctypes.c_uint64
        L.indep_count.argtypes = [ctypes.c_void_p]
        L.indep_copy.argtypes = [ctypes.c_void_p] * 7
        L.indep_free.argtypes = [ctypes.c_void_p]
        L.indep_eval.restype = ctypes.c_double
        L.indep_eval.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_int]
        _INDEP = L
    return _INDEP


@dataclass
class IndepCells:
    """Cells of a z-range that the reference's rule, evaluated WITHOUT the engine's math (f64 + libm or
    f32 + libm), makes active, plus every cell with a corner closer to the surface than `list_below`."""
    keys: np.ndarray      # (n,) u64, label keys, ascending
    active: np.ndarray    # (n,) bool
    nibbles: np.ndarray   # (n,) u8
    positions: np.ndarray  # (n,3) f64 (zeros where not active)
    min_abs: np.ndarray   # (n,) f64: min |corner value| of the cell, SDF units
    corners: np.ndarray   # (n,8) f64
    voxel: float          # min voxel edge length, SDF units
    seconds: float


def indep_run(sdf, res, bounds=2.0, z_begin=0, z_end=None, label_add=1, list_below_voxels=1e-3, precision=64, threads=0,
              bmin=None, bmax=None) -> IndepCells:
    import time
    L = indep_lib()
    sid = SDF_IDS[sdf] if isinstance(sdf, str) else int(sdf)
    res3 = np.array([res] * 3 if np.isscalar(res) else res, dtype=np.uint32)
    if bmin is None:
        bmin, bmax = cube_bounds(bounds)
    bmin = np.ascontiguousarray(bmin, np.float32)
    bmax = np.ascontiguousarray(bmax, np.float32)
    if z_end is None:
        z_end = int(res3[2]) - 1
    voxel = float(np.min((bmax - bmin) / (res3.astype(np.float32) - np.float32(1))))
    t0 = time.perf_counter()
    h = L.indep_run(sid, res3.ctypes.data, bmin.ctypes.data, bmax.ctypes.data, int(z_begin), int(z_end), int(label_add),
                    float(list_below_voxels) * voxel, int(precision), int(threads))
    n = L.indep_count(h)
    keys = np.empty(n, np.uint64); act = np.empty(n, np.uint8); nib = np.empty(n, np.uint8)
    pos = np.empty((n, 3), np.float64); mabs = np.empty(n, np.float64); corners = np.empty((n, 8), np.float64)
    L.indep_copy(h, keys.ctypes.data, act.ctypes.data, nib.ctypes.data, pos.ctypes.data, mabs.ctypes.data, corners.ctypes.data)
    L.indep_free(h)
    return IndepCells(keys, act.astype(bool), nib, pos, mabs, corners, voxel, time.perf_counter() - t0)


def indep_eval(sdf, x, y, z, precision=64) -> float:
    sid = SDF_IDS[sdf] if isinstance(sdf, str) else int(sdf)
    return indep_lib().indep_eval(sid, float(x), float(y), float(z), int(precision))
