/* sdf2mesh_b200.h -- C ABI of libsdf2mesh_b200.so: the B200-native replacement for the
 * SDF -> dual-contoured quad mesh path of WilstonOreo/sdf2mesh.
 *
 * The reference has no plugin/FFI seam: the path is inlined in `run()`
 * (/root/reference/src/bin/sdf2mesh/main.rs:177-364).  This header is the seam a Rust `extern "C"`
 * block, the C++ CLI (cli/sdf2mesh.cpp) and the Python host (sdf2mesh_b200/)
 * all bind.  Every entry point names the reference interface it replaces.
 *
 * Conventions: every call returns an s2m_status (0 = ok); on failure s2m_last_error() returns a
 * thread-local human-readable message (NVRTC log, CUDA error, parse error with line:col).  Handles
 * are opaque.  One host thread drives one s2m_ctx at a time; at most one begin()..finish() pair is
 * outstanding per ctx.  There is no CPU fallback: without a CUDA device s2m_ctx_create fails.
 * Plain pointers and sizes only; no C++/torch types.
 */
#ifndef SDF2MESH_B200_H_
#define SDF2MESH_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum s2m_status {
  S2M_OK = 0,
  S2M_ERR_INVALID_ARG = 1,
  S2M_ERR_IO = 2,
  S2M_ERR_PARSE = 3,       /* reference: ShaderProcessingError::ParseErrors  (shadertoy.rs:70-80) */
  S2M_ERR_VALIDATION = 4,  /* reference: ShaderProcessingError::ValidationError / wgpu validation panic */
  S2M_ERR_MISSING_SDF = 5, /* reference: ShaderProcessingError::MissingSdf(name) */
  S2M_ERR_SHADER = 6,      /* reference: ShaderProcessingError::ShaderError(msg) */
  S2M_ERR_NVRTC = 7,       /* reference: driver shader compile failure inside create_shader_module */
  S2M_ERR_CUDA = 8,
  S2M_ERR_NO_DEVICE = 9,
  S2M_ERR_OOM = 10,
  S2M_ERR_UNSUPPORTED = 11,
  S2M_ERR_STATE = 12
} s2m_status;

const char* s2m_last_error(void);
const char* s2m_version(void);
void s2m_free(void* p); /* frees strings returned through char** out-parameters */

/* ------------------------------------------------------------------ shader source (host only)
 * s2m_shader mirrors `Sdf3DShader` (/root/reference/src/shader.rs:35-40): an opaque holder of ONE
 * assembled source string plus how to lower it. */
typedef struct s2m_shader s2m_shader;

typedef enum s2m_source_kind {
  S2M_SRC_SDF3D = 0,         /* text of a .sdf3d file: WGSL + `use`/`include` lines (shader.rs:159-203) */
  S2M_SRC_GLSL_FRAGMENT = 1, /* GLSL fragment shader containing `float <sdf>(vec3)` (shader.rs:73-104) */
  S2M_SRC_WGSL = 2,          /* already-assembled WGSL (no directive processing) */
  S2M_SRC_CUDA = 3           /* CUDA C++ defining `float sdf3d(vec3 p)` inside namespace s2m_user (diagnostic) */
} s2m_source_kind;

/* Sdf3DShader::from_path (shader.rs:44): never fails on IO; an unreadable file is reported through
 * s2m_shader_log() and yields an empty source, like the reference's log::error!. */
int s2m_shader_from_path(const char* path, s2m_shader** out);
/* Sdf3DShader::from_glsl_fragment_shader (shader.rs:73): S2M_ERR_PARSE / S2M_ERR_MISSING_SDF / S2M_ERR_SHADER. */
int s2m_shader_from_glsl_fragment_shader(const char* path, const char* sdf_name, s2m_shader** out);
/* Same two constructors from memory.  include_dir: directory `include "f";` resolves against
 * (NULL = current working directory, as the reference does). */
int s2m_shader_from_source(const char* text, size_t len, int kind, const char* sdf_name,
                           const char* include_dir, s2m_shader** out);
/* Sdf3DShader::from_shadertoy_api (shader.rs:110-144) without the REST fetch (no network here):
 * `code` is the GLSL of the shader's image pass as ShaderToy serves it (mainImage + helpers); it is
 * wrapped with the ShaderToy uniform block and an empty main() (shadertoy.rs:141-167), converted,
 * and main_1 / main / mainImage are removed.  Same errors as the GLSL constructor. */
int s2m_shader_from_shadertoy_source(const char* code, size_t len, const char* sdf_name, s2m_shader** out);
/* The same from the body of the API response the host fetched (shadertoy.rs:126-131:
 * GET https://www.shadertoy.com/api/v1/shaders/{id}?key=...): {"Shader": {"info": ..., "renderpass": [{"code": ...}]}}
 * -- the code of all passes is concatenated like fetch_code_from_last_pass (:126-132 of impl Shader) -- or
 * {"Error": "..."} -> S2M_ERR_SHADER with the API's message (ShaderProcessingError::ShaderError). */
int s2m_shader_from_shadertoy_response(const char* body, size_t len, const char* sdf_name, s2m_shader** out);
/* Sdf3DShader::add_to_source (shader.rs:155) */
int s2m_shader_add_to_source(s2m_shader* s, const char* text);
/* the `source` field: assembled WGSL (for GLSL input: WGSL regenerated from the IR) */
const char* s2m_shader_source(const s2m_shader* s);
/* Sdf3DShader::write_to_file (shader.rs:206); what --debug-wgsl writes (main.rs:223-225) */
int s2m_shader_write_to_file(const s2m_shader* s, const char* path);
/* info/warn/error lines the reference would have sent to `log` (module names, include errors) */
const char* s2m_shader_log(const s2m_shader* s);
/* lower to CUDA C++ without compiling (returns malloc'd text; S2M_ERR_PARSE/VALIDATION/MISSING_SDF) */
int s2m_shader_lower_to_cuda(const s2m_shader* s, char** cuda_out);
/* the same code over packed f32x2 pairs (two evaluations per call, sm_100a FFMA2; csrc/s2m_pvec.h): what K1
 * compiles next to the scalar form.  An empty string if the shader uses something the packed types do not
 * cover (matrices) -- K1 then evaluates one corner at a time. */
int s2m_shader_lower_to_cuda_packed(const s2m_shader* s, char** cuda_out);
void s2m_shader_free(s2m_shader* s);

/* WGSL text munging used by the GLSL / ShaderToy path (shadertoy.rs:169-352); results malloc'd. */
int s2m_glsl_to_wgsl(const char* glsl, char** wgsl_out);                                /* convert_glsl_to_wgsl :169 */
int s2m_wgsl_remove_function(const char* wgsl, const char* fn_prefix, char** out);      /* :251 */
int s2m_wgsl_has_function(const char* wgsl, const char* fn_name, int* found);           /* :208, :295 */
int s2m_wgsl_rename_function(const char* wgsl, const char* old_name, const char* new_name, char** out); /* :316 */

/* ------------------------------------------------------------------ device context */
typedef struct s2m_ctx s2m_ctx;
/* replaces wgpu Instance/Adapter/Device/Queue setup (main.rs:180-196) */
int s2m_ctx_create(int device_ordinal, s2m_ctx** out);
void s2m_ctx_destroy(s2m_ctx* ctx);
int s2m_ctx_device_info(const s2m_ctx* ctx, char* name, size_t name_len, int* sm_count, uint64_t* total_mem);

/* ------------------------------------------------------------------ module = compiled SDF
 * replaces Sdf3DShader::create_shader_module (shader.rs:220) + create_compute_pipeline (main.rs:283-290):
 * front-end -> CUDA C++ -> NVRTC (sm_100a, --fmad=false) -> cubin -> cuModuleLoadData.
 * ctx may be NULL: compile to cubin only (no device needed; used by CPU-side tests).
 * Cubins are kept on disk -- in $S2M_CACHE_DIR if set (set but empty: no cache), else $XDG_CACHE_HOME/sdf2mesh_b200, else
 * ~/.cache/sdf2mesh_b200 -- one file per distinct translation unit + options + NVRTC version; a repeated compile, also by
 * another process, is a file read. */
typedef struct s2m_module s2m_module;
#define S2M_COMPILE_ALLOW_FMA 1u /* let ptxas contract a*b+c (faster, NOT bit-identical to the oracle) */
int s2m_module_compile(s2m_ctx* ctx, const s2m_shader* shader, uint32_t flags, s2m_module** out);
/* Loads the cubins of an already compiled module (compiled with ctx NULL, or for another device) into
 * ctx as a new module: compile once while the contexts are still being created, then instantiate per
 * GPU.  `compiled` stays valid and is freed separately. */
int s2m_module_instantiate(const s2m_module* compiled, s2m_ctx* ctx, s2m_module** out);
const char* s2m_module_log(const s2m_module* m);         /* NVRTC log */
const char* s2m_module_cuda_source(const s2m_module* m); /* what NVRTC compiled */
int s2m_module_cubin(const s2m_module* m, const void** data, size_t* size); /* part 0 */
/* The module is compiled as up to three NVRTC programs on concurrent host threads: part 0 holds K1
 * (s2m_k1_slab), part 1 K4a (s2m_k4_vertices), part 2 the diagnostic kernels.  S2M_ERR_INVALID_ARG past
 * the last part (S2M_JIT_SPLIT=0 in the environment: one part with everything). */
int s2m_module_cubin_part(const s2m_module* m, int part, const void** data, size_t* size);
double s2m_module_compile_ms(const s2m_module* m, int which); /* 0 front-end, 1 NVRTC, 2 load */
void s2m_module_free(s2m_module* m);

/* ------------------------------------------------------------------ meshing
 * s2m_mesh_params mirrors the AppState uniform (main.rs:28-33; dualcontour.wgsl:133-141):
 * bb_min.xyz, bb_max.xyz, eps (= bb_max.w), dims.xyz.  dims.w (z_slice_idx) disappears: there is
 * no per-slice loop. */
#define S2M_MESH_ALL_SLICES 1u      /* scan all res_z slices, label = z (default reproduces the reference's
                                       one-slice readback lag: slices 0..res_z-2, label = z+1) */
#define S2M_MESH_NO_NORMALS 2u      /* skip sdf3d_normal (normals only reach the PLY writer) */
#define S2M_MESH_EXACT_DENSE 4u     /* reference-cost mode: every cell is a candidate (8 evaluations per cell) */
#define S2M_MESH_KEEP_CANDIDATES 8u /* keep the candidate key list in the result (tests) */
#define S2M_MESH_KEEP_INVALID 32u    /* keep the list of invalid quads (which cell, which edge, which corner is missing) */
#define S2M_MESH_CONSISTENT_CORNERS 64u /* not the reference's arithmetic: a cell's max corner is the next cell's min corner
                                          (bmin + size*(i+1) instead of min + size), so neighbouring cells agree on every
                                          corner value and no quad is lost to 1-ulp disagreements.  With S2M_MESH_ALL_SLICES this is
                                          the watertight mode (CLI --watertight). */
#define S2M_MESH_QUADS_U32 128u      /* quad indices as u32 -- the reference's own type (lib.rs Quad(u32, ..)) -- in
                                       s2m_result_info.quads32 instead of u64 in .quads: half the bytes to copy and keep.
                                       s2m_mesh_finish fails with S2M_ERR_UNSUPPORTED if an index would not fit. */
#define S2M_MESH_CLASSIFY_FROM_SLAB 16u /* K2 re-reads the f32 slab through shared memory instead of K1's class bit planes */
#define S2M_MESH_NO_SLAB 256u        /* slab-free form: K1 writes only the 2-bit corner classes (0.25 B per corner instead of 4.25), K4a
                                       evaluates all 8 corners of every candidate cell.  The default for SDFs that are cheap to evaluate
                                       (s2m_module_prefers_no_slab); same results either way. */
#define S2M_MESH_TIMINGS 1024u        /* fill the per-kernel fields of s2m_timings (k1_slab_ms ... d2h_ms) from CUDA-event spans around every
                                       launch.  Off by default: the records keep consecutive kernels from overlapping head and tail and
                                       reading them back costs ~0.35 ms per run; device_ms / total_ms / host_wall_ms are always measured. */
#define S2M_MESH_RELATIVE_QUADS 512u /* quads keep the slab-relative indices the device wrote: global index = value + quad_index_add
                                       (s2m_result_info; wrapping in the index width -- a vertex of the halo slice is "negative").
                                       s2m_mesh_finish then only waits for the copies instead of adding the base on the host; the
                                       writers (s2m_write_mesh_parts, s2m_result_write_*) accept both forms. */

typedef struct s2m_mesh_params {
  uint32_t struct_size; /* sizeof(s2m_mesh_params) */
  float bb_min[3];
  float bb_max[3];
  float eps;
  uint32_t dims[3];
  uint32_t flags;
  uint32_t z_begin, z_end;     /* z-slab of true cell slices [z_begin, z_end); 0,0 = whole grid */
  float tau_voxels;            /* candidate band half-width in voxels; 0 = default: max(1/16, 443 x the 1-ulp coordinate
                                  mismatch between the reference's corners and the slab's, in voxels) */
  uint64_t slab_budget_bytes;  /* max bytes of corner slab resident at once; 0 = default */
} s2m_mesh_params;

/* AppState::from(&Arguments) (main.rs:139-175): resolution 0 -> 256, rounded UP to a power of two
 * (*rounded = 1 if it was changed); bounds <= 0 or NaN -> 2; cube centred on the origin
 * (lib.rs:115-118); eps = 1e-4. */
int s2m_params_from_cli(uint32_t resolution, float bounds, s2m_mesh_params* out, int* rounded);

typedef struct s2m_timings {
  float k1_slab_ms, k2_classify_ms, k3_compact_ms, k4_vertices_ms, k4_quads_ms;
  float d2h_ms;        /* device->pinned host copies not hidden behind kernels */
  float device_ms;     /* first launch -> last kernel finished, results resident in HBM (CUDA events) */
  float total_ms;      /* first launch -> everything resident in pinned host memory (CUDA events) */
  double host_wall_ms; /* same span by the host clock */
  uint32_t launches;   /* kernel launches issued */
  uint32_t chunks;     /* slab chunks */
} s2m_timings;

typedef struct s2m_result s2m_result;
typedef struct s2m_result_info {
  uint64_t n_vertices;       /* own vertices (halo slice excluded) */
  uint64_t n_halo_vertices;  /* vertices of slice z_begin-1, recomputed locally to name them in quads */
  uint64_t n_quads;          /* valid quads */
  uint64_t n_invalid_quads;  /* reference: "Invalid quad" warnings (mesh.rs:270-278) */
  uint64_t n_candidates;
  /* library-owned pinned host memory, valid until s2m_result_free */
  const float* positions;      /* 3 * n_vertices */
  const float* normals;        /* 3 * n_vertices (zeros with S2M_MESH_NO_NORMALS) */
  const uint64_t* cell_keys;   /* n_vertices: x | y<<16 | label<<32  (mesh.rs:224-226) */
  const uint8_t* sign_nibbles; /* n_vertices: bit0 s100, bit1 s010, bit2 s001, bit3 s000 (main.rs:338-339) */
  const uint64_t* quads;       /* 4 * n_quads global vertex indices, after Quad::swap, reference order */
  const uint64_t* candidates;  /* n_candidates keys (true z) if S2M_MESH_KEEP_CANDIDATES, else NULL */
  /* S2M_MESH_KEEP_INVALID: the quads the reference reports as "Invalid quad: Quad(..)" (mesh.rs:270-278),
   * 6 u64 per record: cell key, edge (0 = X, 1 = Y, 2 = Z), q0..q3 after Quad::swap with
   * UINT64_MAX for a missing vertex; sorted by (key, edge) = the reference's order.  At most 2^20. */
  const uint64_t* invalid_records;
  uint64_t n_invalid_records;
  const float* halo_positions;   /* 3 * n_halo_vertices: positions of the recomputed slice below this slab (vertex
                                    global_vertex_base - n_halo_vertices + j), so a slab can be written on its own */
  int64_t global_vertex_base;    /* what s2m_mesh_finish was given (0 for s2m_mesh_run) */
  const uint32_t* quads32;       /* S2M_MESH_QUADS_U32: 4 * n_quads u32 indices, and .quads is NULL */
  int64_t quad_index_add;        /* global vertex index = quad value + quad_index_add, wrapping in the index width: 0 unless
                                    S2M_MESH_RELATIVE_QUADS (then = global_vertex_base) */
  s2m_timings timings;
} s2m_result_info;

/* replaces main.rs:298-356 (slice loop) + VertexList (mesh.rs:229-265) + VertexList::fetch_triangle_indices
 * (mesh.rs:267-324): K1 slab, K2 classify, K3 compact, K4a vertices, K4b quads, z-chunk by z-chunk, each chunk's
 * vertices and quads copied into pinned host memory while the next chunk computes.  Returns when the last copy has
 * been QUEUED; the vertex count (s2m_result_get) is final, the arrays are not yet.  Quads are written with
 * slab-relative indices (local vertex index, the recomputed halo slice counting as negative), so nothing waits for
 * the global vertex base.
 * Limits: dims in [2, 65535] per axis (cell coordinates are u16 in the reference's key, mesh.rs:214); fewer than 2^32 - 1
 * candidate cells PER SLAB (ranks are u32: S2M_ERR_UNSUPPORTED beyond -- split the grid with z_begin / z_end; 4096^3 of
 * the mandelbulb has 4 x 10^7); vertex and quad indices are 64-bit (u32 on request, checked in s2m_mesh_finish). */
int s2m_mesh_begin(s2m_ctx* ctx, s2m_module* m, const s2m_mesh_params* p, s2m_result** out);
/* Waits for the copies and fixes the slab's place in the whole mesh: global_vertex_base = the exclusive prefix of
 * n_vertices over lower z-slabs (from an all-gather across ranks; 0 on one GPU).  By default the quads in host
 * memory become global indices (+ base, a pass over the quads on a few host threads); with S2M_MESH_RELATIVE_QUADS
 * they stay relative and the base is reported as s2m_result_info.quad_index_add. */
int s2m_mesh_finish(s2m_result* r, int64_t global_vertex_base);
/* begin + finish(0) */
int s2m_mesh_run(s2m_ctx* ctx, s2m_module* m, const s2m_mesh_params* p, s2m_result** out);
int s2m_result_get(const s2m_result* r, s2m_result_info* out);
void s2m_result_free(s2m_result* r);

/* TriangleMesh::write_to_file (mesh.rs:182): by extension, case-insensitive: .stl -> ASCII STL
 * (mesh.rs:167), .ply -> ASCII PLY (mesh.rs:198); unknown extension logs an error and returns OK
 * like the reference.  A z-slab result (s2m_mesh_begin/finish with a halo) can be written as STL on
 * its own (it carries the halo positions its quads refer to); PLY needs the whole mesh. */
int s2m_result_write_mesh(const s2m_result* r, const char* path);
int s2m_result_write_stl_binary(const s2m_result* r, const char* path);
/* The same writers over several z-slab results held by ONE process (e.g. one host thread per GPU),
 * given in z order with consecutive global vertex bases: one STL / PLY file for the whole mesh.
 * binary_stl != 0 writes binary STL for a .stl path. */
int s2m_write_mesh_parts(const s2m_result* const* parts, int n_parts, const char* path, int binary_stl);
/* The same writers over caller-owned host arrays -- TriangleMesh::write_to_file (mesh.rs:182) for a mesh
 * that is not held in an s2m_result (another process's z-slabs, a mesh read back from disk).  Of each
 * s2m_result_info only n_vertices, positions, normals (PLY), n_quads, quads or quads32,
 * global_vertex_base, n_halo_vertices and halo_positions are read.  Needs no device. */
int s2m_write_mesh_arrays(const s2m_result_info* parts, int n_parts, const char* path, int binary_stl);

/* Reads n (<= 32) 64-bit words that live in device memory -- the output of the count all-gather --
 * into host memory with a one-warp kernel writing through mapped pinned memory, on the caller's CUDA
 * stream (cudaStream_t passed as void*; NULL = the ctx's own stream), then waits for that stream.
 * A cudaMemcpy of these few bytes would queue on the copy engine behind the slab's vertex copy
 * (measured: ~2 ms per step on 8 GPUs). */
int s2m_read_device_words(s2m_ctx* ctx, const void* device_words, uint32_t n, uint64_t* out, void* cuda_stream);

/* ------------------------------------------------------------------ several GPUs in one process: z-slabs
 * The reference drives one device from one thread (main.rs:180-196, :298-356).  s2m_multi holds one s2m_ctx per device
 * ordinal, one host thread per device (kept alive between runs) and one NCCL communicator per device (ncclCommInitAll;
 * libnccl.so.2 is loaded with dlopen).  s2m_multi_mesh_run cuts the grid into contiguous z-slabs balanced by cost
 * (s2m_cost_probe on the first device, refined from the measured times of the second and the fourth run on the same SDF and
 * grid), meshes every slab with s2m_mesh_begin on its GPU -- each recomputing the slice below its slab as halo --
 * exchanges the per-slab vertex counts with ONE ncclAllGather (8 bytes per rank) and calls s2m_mesh_finish with the
 * exclusive prefix.  parts_out receives n results in z order (S2M_MESH_RELATIVE_QUADS form: global index = quad value +
 * quad_index_add); free each with s2m_result_free, write them with s2m_write_mesh_parts.  One host thread drives one
 * s2m_multi at a time. */
typedef struct s2m_multi s2m_multi;
#define S2M_MULTI_NO_NCCL 1u       /* exchange the counts through host memory (a barrier between the host threads); only then may a
                                      device ordinal appear more than once (several slabs sharing one GPU: one-GPU test boxes) */
#define S2M_MULTI_EQUAL_SLABS 2u   /* equal-thickness slabs: no cost probe, no refinement */
#define S2M_MULTI_NO_REBALANCE 4u  /* keep the cost-probe partition: no refinement from measured times */
typedef struct s2m_multi_timings {
  int n;                     /* devices */
  double wall_ms;            /* the whole call after partitioning, host clock */
  double begin_ms[64];       /* per device, host clock: s2m_mesh_begin (K1 .. K4b, copies queued) */
  double exchange_ms[64];    /* the count all-gather, including the wait for the slowest device's begin */
  double finish_ms[64];      /* s2m_mesh_finish: the wait for this device's last copies */
} s2m_multi_timings;
int s2m_multi_create(const int* device_ordinals, int n, uint32_t flags, s2m_multi** out);
void s2m_multi_destroy(s2m_multi* mc);
int s2m_multi_size(const s2m_multi* mc);
s2m_ctx* s2m_multi_ctx(s2m_multi* mc, int k);                  /* the k-th device's context (owned by mc) */
int s2m_multi_uses_nccl(const s2m_multi* mc, int* nccl_version); /* 1 if the counts travel through ncclAllGather */
/* `compiled`: a module from s2m_module_compile (with any ctx, or NULL); its cubins are loaded once per device. */
int s2m_multi_mesh_run(s2m_multi* mc, const s2m_module* compiled, const s2m_mesh_params* p, s2m_result** parts_out);
int s2m_multi_get_partition(const s2m_multi* mc, uint32_t* bounds_out /* n + 1 */);
int s2m_multi_last_timings(const s2m_multi* mc, s2m_multi_timings* out);
/* Slab boundaries (host arithmetic, needs no device): b[0..world] over n_slices z-slices, strictly increasing, slab g
 * getting ~1/world of `cost` (relative cost of n_cost equal-thickness z bands; NULL = equal thickness). */
int s2m_partition_slices(uint32_t n_slices, int world, const double* cost, int n_cost, uint32_t* bounds_out);
/* the same boundaries refined from the seconds every slab actually took (cost: the band profile, or NULL) */
int s2m_rebalance_slices(const uint32_t* bounds, int world, const double* seconds, const double* cost, int n_cost, uint32_t* bounds_out);

/* diagnostics */
int s2m_eval_points(s2m_ctx* ctx, s2m_module* m, const float* xyz, uint64_t n, float* out);
/* 1 if K1 of this module evaluates corner pairs in packed f32x2 arithmetic (csrc/s2m_pvec.h) */
int s2m_module_is_packed(const s2m_module* m);
/* a number no other module of this process has (an instance made by s2m_module_instantiate keeps the number of the module
 * it was made from): caches keyed by it survive a module being freed and another one landing at the same address */
uint64_t s2m_module_uid(const s2m_module* m);
/* 1 if this module's SDF is cheap enough that meshing defaults to the slab-free form (S2M_MESH_NO_SLAB) */
int s2m_module_prefers_no_slab(const s2m_module* m);
/* FP32 FMA throughput of the device in TFLOP/s, measured with dependent-chain kernels (no memory traffic, ~2 ms each, best of 3):
 * out[0] FFMA reg,reg,reg; out[1] FFMA reg,imm,imm; out[2] FFMA2 (packed f32x2) pair,bcast,imm.  The denominator of K1's
 * FP32 roofline (bench.py); 148 SMs x 128 lanes x 2 x clock is the nominal figure beside it. */
int s2m_measure_fp32_peak(s2m_ctx* ctx, double out_tflops[3]);
/* the packed form, raw: lane lo evaluates xyz_a[i], lane hi xyz_b[i]; disagreed[i] != 0: the lanes took
 * different decisions and out_b[i] is not valid (K1 re-evaluates such corners on their own).
 * S2M_ERR_UNSUPPORTED if the module has no packed form. */
int s2m_eval_pairs(s2m_ctx* ctx, s2m_module* m, const float* xyz_a, const float* xyz_b, uint64_t n, float* out_a, float* out_b,
                   uint8_t* disagreed);
/* one corner plane of K1's slab: (dims[1]+1) x (dims[0]+1) floats, row-major */
int s2m_debug_slab_plane(s2m_ctx* ctx, s2m_module* m, const s2m_mesh_params* p, uint32_t plane, float* out);
/* relative K1 cost of `planes` equal-thickness z bands (for balancing z-slabs across GPUs): the module's K1 on the middle
 * corner plane of every band, timed with events (S2M_COST_PROBE=lattice: the round-1 lattice of scalar evaluations) */
int s2m_cost_probe(s2m_ctx* ctx, s2m_module* m, const s2m_mesh_params* p, uint32_t planes, double* cost_out);

#ifdef __cplusplus
}
#endif
#endif /* SDF2MESH_B200_H_ */
