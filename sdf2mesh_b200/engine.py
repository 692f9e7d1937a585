"""Context / Module / meshing calls over the C ABI (the path main.rs:177-364 inlines in `run()`)."""
import ctypes
from dataclasses import dataclass

import numpy as np

from . import _capi
from ._capi import MeshParams, ResultInfo, check, lib


class Context:
    """One CUDA device (replaces wgpu Instance/Adapter/Device/Queue, main.rs:180-196)."""

    def __init__(self, device: int = 0):
        self._h = ctypes.c_void_p()
        check(lib().s2m_ctx_create(device, ctypes.byref(self._h)))
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            lib().s2m_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def device_info(self):
        name = ctypes.create_string_buffer(256)
        sms = ctypes.c_int()
        mem = ctypes.c_uint64()
        check(lib().s2m_ctx_device_info(self._h, name, 256, ctypes.byref(sms), ctypes.byref(mem)))
        return name.value.decode(), sms.value, mem.value

    def measure_fp32_peak(self):
        """FP32 FMA throughput in TFLOP/s: (FFMA reg,reg,reg; FFMA reg,imm,imm; FFMA2 pair,bcast,imm)"""
        out = (ctypes.c_double * 3)()
        check(lib().s2m_measure_fp32_peak(self._h, out))
        return tuple(float(v) for v in out)

    def read_device_words(self, device_ptr: int, n: int, cuda_stream: int = 0):
        """n (<= 32) u64 words from device memory, through a kernel + mapped pinned memory on
        `cuda_stream` (a cudaStream_t handle; 0 = the context's stream) instead of a cudaMemcpy"""
        out = (ctypes.c_uint64 * n)()
        check(lib().s2m_read_device_words(self._h, ctypes.c_void_p(device_ptr), n, out, ctypes.c_void_p(cuda_stream)))
        return [int(v) for v in out]


class Module:
    """A compiled SDF (replaces create_shader_module + create_compute_pipeline)."""

    def __init__(self, shader, ctx: Context = None, flags: int = 0):
        self._h = ctypes.c_void_p()
        self.ctx = ctx
        self._shader = shader
        if shader is not None:
            check(lib().s2m_module_compile(ctx._h if ctx is not None else None, shader._h, flags, ctypes.byref(self._h)))

    def instantiate(self, ctx: Context) -> "Module":
        """this module's cubins loaded into ctx as a new Module (compile once with ctx=None, load per GPU)"""
        m = Module(None, ctx)
        m._shader = self._shader
        check(lib().s2m_module_instantiate(self._h, ctx._h, ctypes.byref(m._h)))
        return m

    def close(self):
        if getattr(self, "_h", None):
            lib().s2m_module_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def uid(self) -> int:
        """s2m_module_uid: unique in this process; an instance keeps the number of the module it was made from"""
        return int(lib().s2m_module_uid(self._h))

    @property
    def log(self) -> str:
        return lib().s2m_module_log(self._h).decode("utf-8", "replace")

    @property
    def cuda_source(self) -> str:
        return lib().s2m_module_cuda_source(self._h).decode("utf-8", "replace")

    @property
    def cubin_size(self) -> int:
        data = ctypes.c_void_p()
        size = ctypes.c_size_t()
        check(lib().s2m_module_cubin(self._h, ctypes.byref(data), ctypes.byref(size)))
        return size.value

    def cubins(self):
        """the module's cubins: K1 | K4a | diagnostic kernels (one with everything under S2M_JIT_SPLIT=0)"""
        out = []
        while True:
            data, size = ctypes.c_void_p(), ctypes.c_size_t()
            if lib().s2m_module_cubin_part(self._h, len(out), ctypes.byref(data), ctypes.byref(size)) != 0:
                return out
            out.append(ctypes.string_at(data, size.value))

    def compile_ms(self):
        return tuple(lib().s2m_module_compile_ms(self._h, i) for i in range(3))

    @property
    def packed(self) -> bool:
        """K1 of this module evaluates corner pairs in packed f32x2 arithmetic (csrc/s2m_pvec.h)"""
        return bool(lib().s2m_module_is_packed(self._h))

    @property
    def prefers_no_slab(self) -> bool:
        """meshing with this module defaults to the slab-free form (MESH_NO_SLAB): its SDF is cheap to evaluate"""
        return bool(lib().s2m_module_prefers_no_slab(self._h))

    def eval_pairs(self, pts_a, pts_b):
        """the packed form, raw: -> (values of pts_a, values of pts_b, lanes-disagreed flags)"""
        a = np.ascontiguousarray(pts_a, np.float32).reshape(-1, 3)
        b = np.ascontiguousarray(pts_b, np.float32).reshape(-1, 3)
        assert a.shape == b.shape
        oa, ob = np.empty(a.shape[0], np.float32), np.empty(a.shape[0], np.float32)
        dv = np.empty(a.shape[0], np.uint8)
        check(lib().s2m_eval_pairs(self.ctx._h, self._h, a.ctypes.data, b.ctypes.data, a.shape[0], oa.ctypes.data, ob.ctypes.data, dv.ctypes.data))
        return oa, ob, dv.astype(bool)

    def eval_points(self, pts) -> np.ndarray:
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 3)
        out = np.empty(pts.shape[0], np.float32)
        check(lib().s2m_eval_points(self.ctx._h, self._h, pts.ctypes.data, pts.shape[0], out.ctypes.data))
        return out


def params_from_cli(resolution=None, bounds=None, flags: int = 0):
    """AppState::from(&Arguments), main.rs:139-175.  Returns (MeshParams, rounded: bool)."""
    p = MeshParams()
    rounded = ctypes.c_int()
    check(lib().s2m_params_from_cli(resolution or 0, float(bounds) if bounds else 0.0, ctypes.byref(p), ctypes.byref(rounded)))
    p.flags = flags
    return p, bool(rounded.value)


def make_params(dims, bb_min, bb_max, eps=1e-4, flags=0, z_begin=0, z_end=0, tau_voxels=0.0, slab_budget_bytes=0):
    p = MeshParams()
    p.struct_size = ctypes.sizeof(MeshParams)
    d = [dims] * 3 if np.isscalar(dims) else list(dims)
    for a in range(3):
        p.bb_min[a] = float(bb_min[a])
        p.bb_max[a] = float(bb_max[a])
        p.dims[a] = int(d[a])
    p.eps = eps
    p.flags = flags
    p.z_begin, p.z_end = z_begin, z_end
    p.tau_voxels = tau_voxels
    p.slab_budget_bytes = slab_budget_bytes
    return p


def _view(ptr, n, dtype, shape=None):
    if n == 0 or not ptr:
        a = np.empty(0, dtype)
    else:
        a = np.ctypeslib.as_array(ptr, shape=(n,)).view(dtype)
    return a.reshape(shape) if shape else a


@dataclass
class MeshData:
    """Numpy views into the library's pinned host memory (valid until MeshResult.free())."""
    positions: np.ndarray
    normals: np.ndarray
    keys: np.ndarray
    nibbles: np.ndarray
    quads: np.ndarray
    candidates: np.ndarray
    invalid_records: np.ndarray  # (n, 6) u64: key, edge, q0..q3 (UINT64_MAX = missing); needs MESH_KEEP_INVALID
    n_invalid_quads: int
    n_halo_vertices: int
    n_candidates: int
    timings: dict
    quad_index_add: int = 0      # MESH_RELATIVE_QUADS: global index = quad value + quad_index_add (wrapping in the index width)

    def global_quads(self) -> np.ndarray:
        """quads as global 64-bit vertex indices, whichever form the result holds (a copy)"""
        q = self.quads
        if q.dtype == np.uint32:
            return (q + np.uint32(self.quad_index_add & 0xFFFFFFFF)).astype(np.uint64)
        return q + np.uint64(self.quad_index_add & 0xFFFFFFFFFFFFFFFF)


class MeshResult:
    def __init__(self, handle, ctx):
        self._h = handle
        self._ctx = ctx  # keep the context alive while views exist

    def finish(self, global_vertex_base: int = 0):
        check(lib().s2m_mesh_finish(self._h, global_vertex_base))
        return self

    def info(self) -> ResultInfo:
        i = ResultInfo()
        check(lib().s2m_result_get(self._h, ctypes.byref(i)))
        return i

    def data(self) -> MeshData:
        i = self.info()
        nv, nq = i.n_vertices, i.n_quads
        t = {f[0]: getattr(i.timings, f[0]) for f in _capi.Timings._fields_}
        return MeshData(
            _view(i.positions, nv * 3, np.float32, (-1, 3)), _view(i.normals, nv * 3, np.float32, (-1, 3)),
            _view(i.cell_keys, nv, np.uint64), _view(i.sign_nibbles, nv, np.uint8),
            _view(i.quads32, nq * 4, np.uint32, (-1, 4)) if i.quads32 else _view(i.quads, nq * 4, np.uint64, (-1, 4)),
            _view(i.candidates, i.n_candidates if i.candidates else 0, np.uint64),
            _view(i.invalid_records, i.n_invalid_records * 6 if i.invalid_records else 0, np.uint64, (-1, 6)),
            i.n_invalid_quads, i.n_halo_vertices, i.n_candidates, t, int(i.quad_index_add))

    def write_mesh(self, path):
        """TriangleMesh::write_to_file (mesh.rs:182): .stl / .ply by extension."""
        check(lib().s2m_result_write_mesh(self._h, str(path).encode()))

    def write_stl_binary(self, path):
        check(lib().s2m_result_write_stl_binary(self._h, str(path).encode()))

    def free(self):
        if getattr(self, "_h", None):
            lib().s2m_result_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def write_mesh_parts(parts, path, binary_stl: bool = False):
    """one STL / PLY file from several z-slab results held by this process, in z order"""
    arr = (ctypes.c_void_p * len(parts))(*[p._h for p in parts])
    check(lib().s2m_write_mesh_parts(arr, len(parts), str(path).encode(), 1 if binary_stl else 0))


def write_mesh_arrays(parts, path, binary_stl: bool = False):
    """TriangleMesh::write_to_file (mesh.rs:182) over caller-owned arrays; needs no device.

    parts: [(positions (n,3) f32, normals (n,3) f32 or None, quads (m,4) u64 or u32)] for the whole mesh,
    or one (positions, normals, quads, global_vertex_base, halo_positions) tuple per z-slab in z order."""
    infos = (ResultInfo * len(parts))()
    keep = []
    for k, part in enumerate(parts):
        pos, nrm, quads = part[0], part[1], part[2]
        base = int(part[3]) if len(part) > 3 else 0
        halo = part[4] if len(part) > 4 else None
        pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 3)
        quads = np.ascontiguousarray(quads)
        if quads.dtype != np.uint32:
            quads = quads.astype(np.uint64, copy=False)
        quads = quads.reshape(-1, 4)
        keep += [pos, quads]
        i = infos[k]
        i.n_vertices, i.n_quads, i.global_vertex_base = pos.shape[0], quads.shape[0], base
        i.positions = pos.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
        if nrm is not None:
            nrm = np.ascontiguousarray(nrm, np.float32).reshape(-1, 3)
            assert nrm.shape == pos.shape
            keep.append(nrm)
            i.normals = nrm.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
        if quads.dtype == np.uint32:
            i.quads32 = quads.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32))
        else:
            i.quads = quads.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))
        if halo is not None:
            halo = np.ascontiguousarray(halo, np.float32).reshape(-1, 3)
            keep.append(halo)
            i.n_halo_vertices = halo.shape[0]
            i.halo_positions = halo.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
    check(lib().s2m_write_mesh_arrays(infos, len(parts), str(path).encode(), 1 if binary_stl else 0))
    del keep


def mesh_begin(ctx: Context, module: Module, params: MeshParams) -> MeshResult:
    h = ctypes.c_void_p()
    check(lib().s2m_mesh_begin(ctx._h, module._h, ctypes.byref(params), ctypes.byref(h)))
    return MeshResult(h, ctx)


def mesh_run(ctx: Context, module: Module, params: MeshParams) -> MeshResult:
    h = ctypes.c_void_p()
    check(lib().s2m_mesh_run(ctx._h, module._h, ctypes.byref(params), ctypes.byref(h)))
    return MeshResult(h, ctx)


def debug_slab_plane(ctx: Context, module: Module, params: MeshParams, plane: int) -> np.ndarray:
    out = np.empty((params.dims[1] + 1, params.dims[0] + 1), np.float32)
    check(lib().s2m_debug_slab_plane(ctx._h, module._h, ctypes.byref(params), plane, out.ctypes.data))
    return out


def cost_probe(ctx: Context, module: Module, params: MeshParams, planes: int) -> np.ndarray:
    out = np.empty(planes, np.float64)
    check(lib().s2m_cost_probe(ctx._h, module._h, ctypes.byref(params), planes, out.ctypes.data))
    return out
