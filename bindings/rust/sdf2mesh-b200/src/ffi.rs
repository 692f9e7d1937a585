//! `extern "C"` declarations for `include/sdf2mesh_b200.h`.  NOT COMPILED in the build image (no rustc);
//! `tests/test_capi.py` checks the Python twin of this file against the header symbol by symbol.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_double, c_float, c_int, c_void};

#[repr(C)] pub struct s2m_shader { _p: [u8; 0] }
#[repr(C)] pub struct s2m_ctx { _p: [u8; 0] }
#[repr(C)] pub struct s2m_module { _p: [u8; 0] }
#[repr(C)] pub struct s2m_result { _p: [u8; 0] }
#[repr(C)] pub struct s2m_multi { _p: [u8; 0] }

pub const S2M_OK: c_int = 0;
pub const S2M_ERR_INVALID_ARG: c_int = 1;
pub const S2M_ERR_IO: c_int = 2;
pub const S2M_ERR_PARSE: c_int = 3; // ShaderProcessingError::ParseErrors
pub const S2M_ERR_VALIDATION: c_int = 4; // ShaderProcessingError::ValidationError
pub const S2M_ERR_MISSING_SDF: c_int = 5; // ShaderProcessingError::MissingSdf
pub const S2M_ERR_SHADER: c_int = 6; // ShaderProcessingError::ShaderError
pub const S2M_ERR_NVRTC: c_int = 7;
pub const S2M_ERR_CUDA: c_int = 8;
pub const S2M_ERR_NO_DEVICE: c_int = 9;
pub const S2M_ERR_OOM: c_int = 10;
pub const S2M_ERR_UNSUPPORTED: c_int = 11;
pub const S2M_ERR_STATE: c_int = 12;

pub const S2M_SRC_SDF3D: c_int = 0;
pub const S2M_SRC_GLSL_FRAGMENT: c_int = 1;
pub const S2M_SRC_WGSL: c_int = 2;
pub const S2M_SRC_CUDA: c_int = 3;

pub const S2M_COMPILE_ALLOW_FMA: u32 = 1;
pub const S2M_MESH_ALL_SLICES: u32 = 1;
pub const S2M_MESH_NO_NORMALS: u32 = 2;
pub const S2M_MESH_EXACT_DENSE: u32 = 4;
pub const S2M_MESH_KEEP_CANDIDATES: u32 = 8;
pub const S2M_MESH_CLASSIFY_FROM_SLAB: u32 = 16;
pub const S2M_MESH_KEEP_INVALID: u32 = 32;
pub const S2M_MESH_CONSISTENT_CORNERS: u32 = 64; // with ALL_SLICES: the watertight mode (not the reference's arithmetic)
pub const S2M_MESH_QUADS_U32: u32 = 128; // indices as u32 in quads32 -- what Quad(u32, u32, u32, u32) wants anyway
pub const S2M_MESH_NO_SLAB: u32 = 256; // slab-free form (the default for cheap SDFs)
pub const S2M_MESH_RELATIVE_QUADS: u32 = 512; // global index = quad value + quad_index_add (wrapping)
pub const S2M_MESH_TIMINGS: u32 = 1024;
pub const S2M_MULTI_NO_NCCL: u32 = 1;
pub const S2M_MULTI_EQUAL_SLABS: u32 = 2;
pub const S2M_MULTI_NO_REBALANCE: u32 = 4;

/// == AppState (main.rs:28-33) minus dims.w
#[repr(C)]
#[derive(Clone, Copy, Default, Debug)]
pub struct s2m_mesh_params {
    pub struct_size: u32,
    pub bb_min: [c_float; 3],
    pub bb_max: [c_float; 3],
    pub eps: c_float,
    pub dims: [u32; 3],
    pub flags: u32,
    pub z_begin: u32,
    pub z_end: u32,
    pub tau_voxels: c_float,
    pub slab_budget_bytes: u64,
}

#[repr(C)]
#[derive(Clone, Copy, Default, Debug)]
pub struct s2m_timings {
    pub k1_slab_ms: c_float,
    pub k2_classify_ms: c_float,
    pub k3_compact_ms: c_float,
    pub k4_vertices_ms: c_float,
    pub k4_quads_ms: c_float,
    pub d2h_ms: c_float,
    pub device_ms: c_float,
    pub total_ms: c_float,
    pub host_wall_ms: c_double,
    pub launches: u32,
    pub chunks: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct s2m_result_info {
    pub n_vertices: u64,
    pub n_halo_vertices: u64,
    pub n_quads: u64,
    pub n_invalid_quads: u64,
    pub n_candidates: u64,
    pub positions: *const c_float,    // 3 * n_vertices, pinned host memory owned by the library
    pub normals: *const c_float,      // 3 * n_vertices
    pub cell_keys: *const u64,        // x | y<<16 | label<<32   (mesh.rs:224-226)
    pub sign_nibbles: *const u8,      // bit0 s100, bit1 s010, bit2 s001, bit3 s000 (main.rs:338-339)
    pub quads: *const u64,            // 4 * n_quads, after Quad::swap, reference order
    pub candidates: *const u64,
    pub invalid_records: *const u64,  // 6 u64 per invalid quad: key, edge, q0..q3 (u64::MAX = missing)
    pub n_invalid_records: u64,
    pub halo_positions: *const c_float,
    pub global_vertex_base: i64,
    pub quads32: *const u32,          // with S2M_MESH_QUADS_U32 (then `quads` is null)
    pub quad_index_add: i64,          // S2M_MESH_RELATIVE_QUADS: global index = quad value + this, wrapping in the index width
    pub timings: s2m_timings,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct s2m_multi_timings {
    pub n: c_int,
    pub wall_ms: c_double,
    pub begin_ms: [c_double; 64],
    pub exchange_ms: [c_double; 64],
    pub finish_ms: [c_double; 64],
}

extern "C" {
    pub fn s2m_version() -> *const c_char;
    pub fn s2m_last_error() -> *const c_char;
    pub fn s2m_free(p: *mut c_void);

    // Sdf3DShader (shader.rs:44, :73, :110, :155, :206)
    pub fn s2m_shader_from_path(path: *const c_char, out: *mut *mut s2m_shader) -> c_int;
    pub fn s2m_shader_from_glsl_fragment_shader(path: *const c_char, sdf: *const c_char, out: *mut *mut s2m_shader) -> c_int;
    pub fn s2m_shader_from_source(text: *const c_char, len: usize, kind: c_int, sdf: *const c_char, include_dir: *const c_char, out: *mut *mut s2m_shader) -> c_int;
    pub fn s2m_shader_from_shadertoy_source(code: *const c_char, len: usize, sdf: *const c_char, out: *mut *mut s2m_shader) -> c_int;
    pub fn s2m_shader_from_shadertoy_response(body: *const c_char, len: usize, sdf: *const c_char, out: *mut *mut s2m_shader) -> c_int;
    pub fn s2m_shader_add_to_source(s: *mut s2m_shader, text: *const c_char) -> c_int;
    pub fn s2m_shader_source(s: *const s2m_shader) -> *const c_char;
    pub fn s2m_shader_write_to_file(s: *const s2m_shader, path: *const c_char) -> c_int;
    pub fn s2m_shader_log(s: *const s2m_shader) -> *const c_char;
    pub fn s2m_shader_lower_to_cuda(s: *const s2m_shader, cuda_out: *mut *mut c_char) -> c_int;
    pub fn s2m_shader_lower_to_cuda_packed(s: *const s2m_shader, cuda_out: *mut *mut c_char) -> c_int;
    pub fn s2m_shader_free(s: *mut s2m_shader);

    // WGSL text munging of the GLSL / ShaderToy path (shadertoy.rs:169-352)
    pub fn s2m_glsl_to_wgsl(glsl: *const c_char, wgsl_out: *mut *mut c_char) -> c_int;
    pub fn s2m_wgsl_remove_function(wgsl: *const c_char, fn_prefix: *const c_char, out: *mut *mut c_char) -> c_int;
    pub fn s2m_wgsl_has_function(wgsl: *const c_char, fn_name: *const c_char, found: *mut c_int) -> c_int;
    pub fn s2m_wgsl_rename_function(wgsl: *const c_char, old_name: *const c_char, new_name: *const c_char, out: *mut *mut c_char) -> c_int;

    // device + module (replaces main.rs:180-196 and shader.rs:220 + main.rs:283-290)
    pub fn s2m_ctx_create(device: c_int, out: *mut *mut s2m_ctx) -> c_int;
    pub fn s2m_ctx_destroy(ctx: *mut s2m_ctx);
    pub fn s2m_ctx_device_info(ctx: *const s2m_ctx, name: *mut c_char, name_len: usize, sm_count: *mut c_int, total_mem: *mut u64) -> c_int;
    pub fn s2m_module_compile(ctx: *mut s2m_ctx, shader: *const s2m_shader, flags: u32, out: *mut *mut s2m_module) -> c_int;
    pub fn s2m_module_log(m: *const s2m_module) -> *const c_char;
    pub fn s2m_module_cuda_source(m: *const s2m_module) -> *const c_char;
    pub fn s2m_module_instantiate(compiled: *const s2m_module, ctx: *mut s2m_ctx, out: *mut *mut s2m_module) -> c_int;
    pub fn s2m_module_cubin(m: *const s2m_module, data: *mut *const c_void, size: *mut usize) -> c_int;
    pub fn s2m_module_cubin_part(m: *const s2m_module, part: c_int, data: *mut *const c_void, size: *mut usize) -> c_int;
    pub fn s2m_module_compile_ms(m: *const s2m_module, which: c_int) -> c_double;
    pub fn s2m_module_is_packed(m: *const s2m_module) -> c_int;
    pub fn s2m_module_uid(m: *const s2m_module) -> u64;
    pub fn s2m_module_prefers_no_slab(m: *const s2m_module) -> c_int;
    pub fn s2m_module_free(m: *mut s2m_module);

    // meshing (replaces main.rs:298-356 and mesh.rs:229-331)
    pub fn s2m_params_from_cli(resolution: u32, bounds: c_float, out: *mut s2m_mesh_params, rounded: *mut c_int) -> c_int;
    pub fn s2m_mesh_begin(ctx: *mut s2m_ctx, m: *mut s2m_module, p: *const s2m_mesh_params, out: *mut *mut s2m_result) -> c_int;
    pub fn s2m_mesh_finish(r: *mut s2m_result, global_vertex_base: i64) -> c_int;
    pub fn s2m_mesh_run(ctx: *mut s2m_ctx, m: *mut s2m_module, p: *const s2m_mesh_params, out: *mut *mut s2m_result) -> c_int;
    pub fn s2m_result_get(r: *const s2m_result, out: *mut s2m_result_info) -> c_int;
    pub fn s2m_result_write_mesh(r: *const s2m_result, path: *const c_char) -> c_int; // mesh.rs:182
    pub fn s2m_result_write_stl_binary(r: *const s2m_result, path: *const c_char) -> c_int;
    pub fn s2m_result_free(r: *mut s2m_result);
    pub fn s2m_write_mesh_arrays(parts: *const s2m_result_info, n_parts: c_int, path: *const c_char, binary_stl: c_int) -> c_int;
    pub fn s2m_write_mesh_parts(parts: *const *const s2m_result, n_parts: c_int, path: *const c_char, binary_stl: c_int) -> c_int;
    pub fn s2m_read_device_words(ctx: *mut s2m_ctx, device_words: *const c_void, n: u32, out: *mut u64, cuda_stream: *mut c_void) -> c_int;

    // several GPUs in one process: z-slabs, one ncclAllGather of the vertex counts (replaces main.rs:177-364 for N devices)
    pub fn s2m_multi_create(device_ordinals: *const c_int, n: c_int, flags: u32, out: *mut *mut s2m_multi) -> c_int;
    pub fn s2m_multi_destroy(mc: *mut s2m_multi);
    pub fn s2m_multi_size(mc: *const s2m_multi) -> c_int;
    pub fn s2m_multi_ctx(mc: *mut s2m_multi, k: c_int) -> *mut s2m_ctx;
    pub fn s2m_multi_uses_nccl(mc: *const s2m_multi, nccl_version: *mut c_int) -> c_int;
    pub fn s2m_multi_mesh_run(mc: *mut s2m_multi, compiled: *const s2m_module, p: *const s2m_mesh_params, parts_out: *mut *mut s2m_result) -> c_int;
    pub fn s2m_multi_get_partition(mc: *const s2m_multi, bounds_out: *mut u32) -> c_int;
    pub fn s2m_multi_last_timings(mc: *const s2m_multi, out: *mut s2m_multi_timings) -> c_int;
    pub fn s2m_partition_slices(n_slices: u32, world: c_int, cost: *const c_double, n_cost: c_int, bounds_out: *mut u32) -> c_int;
    pub fn s2m_rebalance_slices(bounds: *const u32, world: c_int, seconds: *const c_double, cost: *const c_double, n_cost: c_int, bounds_out: *mut u32) -> c_int;

    // diagnostics
    pub fn s2m_measure_fp32_peak(ctx: *mut s2m_ctx, out_tflops: *mut c_double) -> c_int;
    pub fn s2m_eval_points(ctx: *mut s2m_ctx, m: *mut s2m_module, xyz: *const c_float, n: u64, out: *mut c_float) -> c_int;
    pub fn s2m_eval_pairs(ctx: *mut s2m_ctx, m: *mut s2m_module, xyz_a: *const c_float, xyz_b: *const c_float, n: u64,
                          out_a: *mut c_float, out_b: *mut c_float, disagreed: *mut u8) -> c_int;
    pub fn s2m_debug_slab_plane(ctx: *mut s2m_ctx, m: *mut s2m_module, p: *const s2m_mesh_params, plane: u32, out: *mut c_float) -> c_int;
    pub fn s2m_cost_probe(ctx: *mut s2m_ctx, m: *mut s2m_module, p: *const s2m_mesh_params, planes: u32, cost_out: *mut c_double) -> c_int;
}
