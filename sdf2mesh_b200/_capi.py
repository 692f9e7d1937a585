"""ctypes binding of libsdf2mesh_b200.so (include/sdf2mesh_b200.h).

The library is built in-tree by `__graft_entry__.build()` / `make -C sdf2mesh_b200/csrc`.  There is
no CPU fallback: if the shared library is missing the import fails loudly.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsdf2mesh_b200.so")

S2M_OK = 0
STATUS_NAMES = {
    0: "OK", 1: "INVALID_ARG", 2: "IO", 3: "PARSE", 4: "VALIDATION", 5: "MISSING_SDF", 6: "SHADER",
    7: "NVRTC", 8: "CUDA", 9: "NO_DEVICE", 10: "OOM", 11: "UNSUPPORTED", 12: "STATE",
    100: "REQUEST",   # host side only (ShaderProcessingError::RequestError, shadertoy.rs:71): the C library makes no requests
}
ERR_SHADER, ERR_REQUEST = 6, 100
SRC_SDF3D, SRC_GLSL_FRAGMENT, SRC_WGSL, SRC_CUDA = 0, 1, 2, 3
COMPILE_ALLOW_FMA = 1
MESH_ALL_SLICES, MESH_NO_NORMALS, MESH_EXACT_DENSE, MESH_KEEP_CANDIDATES, MESH_CLASSIFY_FROM_SLAB = 1, 2, 4, 8, 16
MESH_KEEP_INVALID = 32
MESH_CONSISTENT_CORNERS = 64
MESH_QUADS_U32 = 128
MESH_NO_SLAB = 256
MESH_RELATIVE_QUADS = 512
MESH_TIMINGS = 1024


class S2mError(RuntimeError):
    """Any non-zero s2m_status.  `.status` is the numeric code, `.kind` its name.

    Mirrors the reference's error surface: ShaderProcessingError variants
    (/root/reference/src/shadertoy.rs:70-80) map to PARSE / VALIDATION / MISSING_SDF / SHADER;
    GPU-side panics (unwrap/expect in main.rs) map to NVRTC / CUDA / NO_DEVICE / OOM.
    """

    def __init__(self, status, message):
        super().__init__(f"[{STATUS_NAMES.get(status, status)}] {message}")
        self.status = status
        self.kind = STATUS_NAMES.get(status, str(status))


class MeshParams(ctypes.Structure):
    _fields_ = [
        ("struct_size", ctypes.c_uint32),
        ("bb_min", ctypes.c_float * 3),
        ("bb_max", ctypes.c_float * 3),
        ("eps", ctypes.c_float),
        ("dims", ctypes.c_uint32 * 3),
        ("flags", ctypes.c_uint32),
        ("z_begin", ctypes.c_uint32),
        ("z_end", ctypes.c_uint32),
        ("tau_voxels", ctypes.c_float),
        ("slab_budget_bytes", ctypes.c_uint64),
    ]


class Timings(ctypes.Structure):
    _fields_ = [
        ("k1_slab_ms", ctypes.c_float), ("k2_classify_ms", ctypes.c_float), ("k3_compact_ms", ctypes.c_float),
        ("k4_vertices_ms", ctypes.c_float), ("k4_quads_ms", ctypes.c_float), ("d2h_ms", ctypes.c_float),
        ("device_ms", ctypes.c_float), ("total_ms", ctypes.c_float), ("host_wall_ms", ctypes.c_double),
        ("launches", ctypes.c_uint32), ("chunks", ctypes.c_uint32),
    ]


class ResultInfo(ctypes.Structure):
    _fields_ = [
        ("n_vertices", ctypes.c_uint64), ("n_halo_vertices", ctypes.c_uint64), ("n_quads", ctypes.c_uint64),
        ("n_invalid_quads", ctypes.c_uint64), ("n_candidates", ctypes.c_uint64),
        ("positions", ctypes.POINTER(ctypes.c_float)), ("normals", ctypes.POINTER(ctypes.c_float)),
        ("cell_keys", ctypes.POINTER(ctypes.c_uint64)), ("sign_nibbles", ctypes.POINTER(ctypes.c_uint8)),
        ("quads", ctypes.POINTER(ctypes.c_uint64)), ("candidates", ctypes.POINTER(ctypes.c_uint64)),
        ("invalid_records", ctypes.POINTER(ctypes.c_uint64)), ("n_invalid_records", ctypes.c_uint64),
        ("halo_positions", ctypes.POINTER(ctypes.c_float)), ("global_vertex_base", ctypes.c_int64),
        ("quads32", ctypes.POINTER(ctypes.c_uint32)), ("quad_index_add", ctypes.c_int64),
        ("timings", Timings),
    ]


class MultiTimings(ctypes.Structure):
    _fields_ = [("n", ctypes.c_int), ("wall_ms", ctypes.c_double), ("begin_ms", ctypes.c_double * 64),
                ("exchange_ms", ctypes.c_double * 64), ("finish_ms", ctypes.c_double * 64)]


MULTI_NO_NCCL, MULTI_EQUAL_SLABS, MULTI_NO_REBALANCE = 1, 2, 4

# every symbol include/sdf2mesh_b200.h declares: name -> (restype, argtypes)
_P = ctypes.c_void_p
_PP = ctypes.POINTER(ctypes.c_void_p)
_S = ctypes.c_char_p
_SP = ctypes.POINTER(ctypes.c_char_p)
SYMBOLS = {
    "s2m_last_error": (_S, []),
    "s2m_version": (_S, []),
    "s2m_free": (None, [_P]),
    "s2m_shader_from_path": (ctypes.c_int, [_S, _PP]),
    "s2m_shader_from_glsl_fragment_shader": (ctypes.c_int, [_S, _S, _PP]),
    "s2m_shader_from_source": (ctypes.c_int, [_S, ctypes.c_size_t, ctypes.c_int, _S, _S, _PP]),
    "s2m_shader_from_shadertoy_source": (ctypes.c_int, [_S, ctypes.c_size_t, _S, _PP]),
    "s2m_shader_from_shadertoy_response": (ctypes.c_int, [_S, ctypes.c_size_t, _S, _PP]),
    "s2m_shader_add_to_source": (ctypes.c_int, [_P, _S]),
    "s2m_shader_source": (_S, [_P]),
    "s2m_shader_write_to_file": (ctypes.c_int, [_P, _S]),
    "s2m_shader_log": (_S, [_P]),
    "s2m_shader_lower_to_cuda": (ctypes.c_int, [_P, _PP]),
    "s2m_shader_lower_to_cuda_packed": (ctypes.c_int, [_P, _PP]),
    "s2m_shader_free": (None, [_P]),
    "s2m_glsl_to_wgsl": (ctypes.c_int, [_S, _PP]),
    "s2m_wgsl_remove_function": (ctypes.c_int, [_S, _S, _PP]),
    "s2m_wgsl_has_function": (ctypes.c_int, [_S, _S, ctypes.POINTER(ctypes.c_int)]),
    "s2m_wgsl_rename_function": (ctypes.c_int, [_S, _S, _S, _PP]),
    "s2m_ctx_create": (ctypes.c_int, [ctypes.c_int, _PP]),
    "s2m_ctx_destroy": (None, [_P]),
    "s2m_ctx_device_info": (ctypes.c_int, [_P, ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_uint64)]),
    "s2m_module_compile": (ctypes.c_int, [_P, _P, ctypes.c_uint32, _PP]),
    "s2m_module_log": (_S, [_P]),
    "s2m_module_cuda_source": (_S, [_P]),
    "s2m_module_instantiate": (ctypes.c_int, [_P, _P, _PP]),
    "s2m_module_cubin": (ctypes.c_int, [_P, _PP, ctypes.POINTER(ctypes.c_size_t)]),
    "s2m_module_cubin_part": (ctypes.c_int, [_P, ctypes.c_int, _PP, ctypes.POINTER(ctypes.c_size_t)]),
    "s2m_module_compile_ms": (ctypes.c_double, [_P, ctypes.c_int]),
    "s2m_module_free": (None, [_P]),
    "s2m_params_from_cli": (ctypes.c_int, [ctypes.c_uint32, ctypes.c_float, ctypes.POINTER(MeshParams), ctypes.POINTER(ctypes.c_int)]),
    "s2m_mesh_begin": (ctypes.c_int, [_P, _P, ctypes.POINTER(MeshParams), _PP]),
    "s2m_mesh_finish": (ctypes.c_int, [_P, ctypes.c_int64]),
    "s2m_mesh_run": (ctypes.c_int, [_P, _P, ctypes.POINTER(MeshParams), _PP]),
    "s2m_result_get": (ctypes.c_int, [_P, ctypes.POINTER(ResultInfo)]),
    "s2m_result_free": (None, [_P]),
    "s2m_result_write_mesh": (ctypes.c_int, [_P, _S]),
    "s2m_result_write_stl_binary": (ctypes.c_int, [_P, _S]),
    "s2m_write_mesh_parts": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, _S, ctypes.c_int]),
    "s2m_write_mesh_arrays": (ctypes.c_int, [_P, ctypes.c_int, _S, ctypes.c_int]),
    "s2m_eval_points": (ctypes.c_int, [_P, _P, _P, ctypes.c_uint64, _P]),
    "s2m_module_is_packed": (ctypes.c_int, [_P]),
    "s2m_module_uid": (ctypes.c_uint64, [_P]),
    "s2m_module_prefers_no_slab": (ctypes.c_int, [_P]),
    "s2m_measure_fp32_peak": (ctypes.c_int, [_P, ctypes.POINTER(ctypes.c_double)]),
    "s2m_eval_pairs": (ctypes.c_int, [_P, _P, _P, _P, ctypes.c_uint64, _P, _P, _P]),
    "s2m_debug_slab_plane": (ctypes.c_int, [_P, _P, ctypes.POINTER(MeshParams), ctypes.c_uint32, _P]),
    "s2m_cost_probe": (ctypes.c_int, [_P, _P, ctypes.POINTER(MeshParams), ctypes.c_uint32, _P]),
    "s2m_read_device_words": (ctypes.c_int, [_P, _P, ctypes.c_uint32, _P, _P]),
    "s2m_multi_create": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.c_uint32, _PP]),
    "s2m_multi_destroy": (None, [_P]),
    "s2m_multi_size": (ctypes.c_int, [_P]),
    "s2m_multi_ctx": (_P, [_P, ctypes.c_int]),
    "s2m_multi_uses_nccl": (ctypes.c_int, [_P, ctypes.POINTER(ctypes.c_int)]),
    "s2m_multi_mesh_run": (ctypes.c_int, [_P, _P, ctypes.POINTER(MeshParams), ctypes.POINTER(ctypes.c_void_p)]),
    "s2m_multi_get_partition": (ctypes.c_int, [_P, ctypes.POINTER(ctypes.c_uint32)]),
    "s2m_multi_last_timings": (ctypes.c_int, [_P, ctypes.POINTER(MultiTimings)]),
    "s2m_partition_slices": (ctypes.c_int, [ctypes.c_uint32, ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.c_int, ctypes.POINTER(ctypes.c_uint32)]),
    "s2m_rebalance_slices": (ctypes.c_int, [ctypes.POINTER(ctypes.c_uint32), ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double), ctypes.c_int, ctypes.POINTER(ctypes.c_uint32)]),
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C sdf2mesh_b200/csrc`).  sdf2mesh_b200 has no CPU fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError if the library does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(status):
    if status != S2M_OK:
        raise S2mError(status, lib().s2m_last_error().decode("utf-8", "replace"))


def take_string(char_pp) -> str:
    """Copy and free a malloc'd char* returned through an out-parameter."""
    s = ctypes.cast(char_pp, ctypes.c_char_p).value
    out = s.decode("utf-8", "replace") if s is not None else ""
    lib().s2m_free(char_pp)
    return out
