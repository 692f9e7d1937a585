// engine.cpp -- device context, NVRTC JIT and the meshing pipeline behind the C ABI
// (include/sdf2mesh_b200.h).  Replaces the wgpu device / pipeline / per-slice texture readback /
// host mesh assembly of the reference:
//   /root/reference/src/bin/sdf2mesh/main.rs:180-196 (device), :229-290 (module, pipeline),
//   :298-356 (slice loop), /root/reference/src/texture.rs (storage textures + MAP_READ staging),
//   /root/reference/src/mesh.rs:213-341 (VertexList -> quads).
//
// Data layout in HBM (all owned by the ctx and reused across runs):
//   slab        f32 [planes][res_y+1][pitch_x]   corner values, variant-A coordinates, chunked in z
//   cand_mask   u32 [nz][res_y][words_x]          1 bit per cell: candidate
//   word_prefix u32 [nz][res_y][words_x]          exclusive popcount prefix (rank table)
//   cand_key    u64 [n_cand]                      x | y<<16 | z_true<<32, linear cell order
//   cand_vrank  u32 [n_cand]                      vertex index or 0xffffffff
//   vert_*      pos f32x3, nrm f32x3, key u64, nibble u8   [n_vertices]
//   quads       u64 [n_quads][4]
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvrtc.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <unistd.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <system_error>
#include <thread>
#include <vector>

#include "common.h"
#include "kernels_static.h"

namespace s2m_internal {
static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
int fail(int status, const std::string& msg) { g_err = msg; return status; }
}  // namespace s2m_internal
using s2m_internal::fail;

extern "C" const char* s2m_last_error(void) { return s2m_internal::g_err.c_str(); }
extern "C" const char* s2m_version(void) { return "sdf2mesh_b200 0.1 (sm_100a)"; }
extern "C" void s2m_free(void* p) { free(p); }

// ------------------------------------------------------------------ driver API (dlopen'ed so the
// library loads on machines without a GPU driver; the front-end and NVRTC work there)
namespace {
typedef int CUresult_t;
typedef struct CUmod_st* CUmodule_t;
typedef struct CUfunc_st* CUfunction_t;
struct DriverApi {
  void* handle = nullptr;
  CUresult_t (*cuInit)(unsigned) = nullptr;
  CUresult_t (*cuModuleLoadData)(CUmodule_t*, const void*) = nullptr;
  CUresult_t (*cuModuleUnload)(CUmodule_t) = nullptr;
  CUresult_t (*cuModuleGetFunction)(CUfunction_t*, CUmodule_t, const char*) = nullptr;
  CUresult_t (*cuLaunchKernel)(CUfunction_t, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
                               cudaStream_t, void**, void**) = nullptr;
  CUresult_t (*cuGetErrorString)(CUresult_t, const char**) = nullptr;
  bool ok = false;
  std::string why;
};
DriverApi& driver() {
  static DriverApi d;
  static std::once_flag once;
  std::call_once(once, [] {
    d.handle = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (!d.handle) { d.why = std::string("cannot load libcuda.so.1: ") + dlerror(); return; }
#define S2M_SYM(name, sym)                                                        \
  d.name = reinterpret_cast<decltype(d.name)>(dlsym(d.handle, sym));              \
  if (!d.name) { d.why = std::string("libcuda.so.1 lacks ") + sym; return; }
    S2M_SYM(cuInit, "cuInit")
    S2M_SYM(cuModuleLoadData, "cuModuleLoadData")
    S2M_SYM(cuModuleUnload, "cuModuleUnload")
    S2M_SYM(cuModuleGetFunction, "cuModuleGetFunction")
    S2M_SYM(cuLaunchKernel, "cuLaunchKernel")
    S2M_SYM(cuGetErrorString, "cuGetErrorString")
#undef S2M_SYM
    d.ok = true;
  });
  return d;
}
std::string cu_err(CUresult_t r) {
  const char* s = nullptr;
  if (driver().cuGetErrorString) driver().cuGetErrorString(r, &s);
  return s ? s : ("CUresult " + std::to_string(r));
}

#define CUDA_TRY(expr)                                                                                   \
  do {                                                                                                   \
    cudaError_t e__ = (expr);                                                                            \
    if (e__ != cudaSuccess)                                                                              \
      return fail(e__ == cudaErrorMemoryAllocation ? S2M_ERR_OOM : S2M_ERR_CUDA,                         \
                  std::string(#expr) + ": " + cudaGetErrorString(e__));                                  \
  } while (0)

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return S2M_OK;
    if (p) { cudaFree(p); p = nullptr; cap = 0; }
    size_t want = bytes + (bytes >> 4) + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      cudaGetLastError();
      e = cudaMalloc(&p, bytes);
      want = bytes;
    }
    if (e != cudaSuccess) { p = nullptr; return fail(S2M_ERR_OOM, "cudaMalloc(" + std::to_string(bytes) + " B): " + cudaGetErrorString(e)); }
    cap = want;
    return S2M_OK;
  }
  // grow, keeping the first `used` bytes (device-to-device copy on `st`)
  int ensure_preserve(size_t bytes, size_t used, cudaStream_t st) {
    if (bytes <= cap) return S2M_OK;
    if (!p || used == 0) return ensure(bytes);
    void* np = nullptr;
    size_t want = bytes + bytes / 2 + 256;
    cudaError_t e = cudaMalloc(&np, want);
    if (e != cudaSuccess) { cudaGetLastError(); want = bytes; e = cudaMalloc(&np, want); }
    if (e != cudaSuccess) return fail(S2M_ERR_OOM, "cudaMalloc(" + std::to_string(bytes) + " B): " + cudaGetErrorString(e));
    e = cudaMemcpyAsync(np, p, used, cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { cudaFree(np); return fail(S2M_ERR_CUDA, std::string("grow copy: ") + cudaGetErrorString(e)); }
    cudaFree(p);
    p = np; cap = want;
    return S2M_OK;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct PinnedBlock { void* p; size_t cap; bool used; };

double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
// S2M_TRACE=1: host-clock timestamps of the pipeline phases on stderr
struct Trace {
  bool on;
  double t0, last;
  Trace() : on(getenv("S2M_TRACE") != nullptr), t0(now_ms()), last(t0) {}
  void mark(const char* what) {
    if (!on) return;
    const double t = now_ms();
    fprintf(stderr, "[s2m trace] %-28s +%8.3f ms (at %8.3f)\n", what, t - last, t - t0);
    last = t;
  }
};
}  // namespace

// ------------------------------------------------------------------ ctx
struct s2m_ctx {
  int device = 0;
  cudaDeviceProp prop{};
  cudaStream_t stream = nullptr, copy_stream = nullptr, prod_stream = nullptr;
  DevBuf slab, cls, slab2, cls2, cand_mask, word_prefix, cand_key, cand_vrank, status, counters;
  DevBuf v_pos, v_nrm, v_key, v_nib, quads, scratch, invalid;
  std::vector<PinnedBlock> pinned;
  unsigned long long* h_counters = nullptr;  // pinned, 16 words
  unsigned long long* h_words = nullptr;     // pinned + mapped, 32 words (s2m_read_device_words)
  cudaEvent_t ev[16]{};
  std::vector<cudaEvent_t> ev_pool;   // per-launch timing events, grown on demand
  uint64_t hint_nv = 0, hint_nq = 0;  // output sizes of the previous run (pinned capacity guess)
  bool busy = false;  // a begin() without finish()/free() is outstanding
  uint32_t publish_launches = 0;  // k_publish launches since the last begin() (they count as kernel launches too)

  void* lease_pinned(size_t bytes) {
    if (bytes == 0) bytes = 16;
    PinnedBlock* best = nullptr;
    for (auto& b : pinned)
      if (!b.used && b.cap >= bytes && (!best || b.cap < best->cap)) best = &b;
    if (best && best->cap <= bytes * 4 + (1u << 20)) { best->used = true; return best->p; }
    void* p = nullptr;
    size_t want = bytes + (bytes >> 3);
    if (cudaHostAlloc(&p, want, cudaHostAllocDefault) != cudaSuccess) {
      cudaGetLastError();
      // drop idle blocks and retry with the exact size
      for (auto it = pinned.begin(); it != pinned.end();)
        if (!it->used) { cudaFreeHost(it->p); it = pinned.erase(it); } else ++it;
      want = bytes;
      if (cudaHostAlloc(&p, want, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    }
    pinned.push_back({p, want, true});
    return p;
  }
  void release_pinned(void* p) {
    for (auto& b : pinned) if (b.p == p) b.used = false;
  }
};

enum Counter { C_NCAND = 0, C_NVERT = 1, C_NHALO = 2, C_NQUAD = 3, C_NINVALID = 4, C_TICKET0 = 5, C_TICKET1 = 6, C_TICKET2 = 7, C_INVALID_CURSOR = 8,
               C_CHUNK_CAND0 = 9, C_CHUNK_CAND1 = 10 /* candidates of the chunk K2 classified into slab buffer 0 / 1 */, C_COUNT = 16 };
enum CtxEvent { EV_BEGIN = 0, EV_KERNELS_DONE = 1, EV_ALL_DONE = 2, EV_PRODUCED0 = 4, EV_PRODUCED1 = 5, EV_CONSUMED0 = 6, EV_CONSUMED1 = 7 };
constexpr unsigned long long kInvalidCapacity = 1ull << 20;

extern "C" int s2m_ctx_create(int device_ordinal, s2m_ctx** out) {
  if (!out) return fail(S2M_ERR_INVALID_ARG, "s2m_ctx_create: out is NULL");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(S2M_ERR_NO_DEVICE, std::string("no CUDA device (this engine has no CPU fallback): ") +
                                       (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
  }
  if (device_ordinal < 0 || device_ordinal >= n) return fail(S2M_ERR_INVALID_ARG, "device ordinal out of range");
  if (!driver().ok) return fail(S2M_ERR_NO_DEVICE, driver().why);
  CUDA_TRY(cudaSetDevice(device_ordinal));
  CUDA_TRY(cudaFree(0));  // create the primary context the driver-API calls will use
  std::unique_ptr<s2m_ctx> c(new s2m_ctx());
  c->device = device_ordinal;
  CUDA_TRY(cudaGetDeviceProperties(&c->prop, device_ordinal));
  if (c->prop.major < 10)
    return fail(S2M_ERR_NO_DEVICE, std::string("device '") + c->prop.name + "' is sm_" + std::to_string(c->prop.major) +
                                       std::to_string(c->prop.minor) + "; this library only carries sm_100a code");
  // Two compute streams: K1/K2 of chunk c+1 (producer, low priority) run while K3/K4a/K4b of chunk c
  // (consumer = the main stream, high priority) execute; see mesh_begin_impl.
  int prio_least = 0, prio_greatest = 0;
  CUDA_TRY(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
  CUDA_TRY(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio_greatest));
  CUDA_TRY(cudaStreamCreateWithPriority(&c->prod_stream, cudaStreamNonBlocking, prio_least));
  CUDA_TRY(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  for (auto& ev : c->ev) CUDA_TRY(cudaEventCreate(&ev));
  CUDA_TRY(cudaHostAlloc(reinterpret_cast<void**>(&c->h_counters), C_COUNT * 8, cudaHostAllocMapped | cudaHostAllocPortable));
  CUDA_TRY(cudaHostAlloc(reinterpret_cast<void**>(&c->h_words), 32 * 8, cudaHostAllocMapped | cudaHostAllocPortable));
  int st = c->counters.ensure(C_COUNT * 8);
  if (st) return st;
  *out = c.release();
  return S2M_OK;
}

extern "C" void s2m_ctx_destroy(s2m_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  for (DevBuf* b : {&c->slab, &c->cls, &c->slab2, &c->cls2, &c->cand_mask, &c->word_prefix, &c->cand_key, &c->cand_vrank, &c->status, &c->counters,
                    &c->v_pos, &c->v_nrm, &c->v_key, &c->v_nib, &c->quads, &c->scratch, &c->invalid})
    b->release();
  for (auto& b : c->pinned) cudaFreeHost(b.p);
  if (c->h_counters) cudaFreeHost(c->h_counters);
  if (c->h_words) cudaFreeHost(c->h_words);
  for (auto& ev : c->ev) if (ev) cudaEventDestroy(ev);
  for (auto& ev : c->ev_pool) if (ev) cudaEventDestroy(ev);
  if (c->stream) cudaStreamDestroy(c->stream);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->prod_stream) cudaStreamDestroy(c->prod_stream);
  delete c;
}

extern "C" int s2m_ctx_device_info(const s2m_ctx* c, char* name, size_t name_len, int* sm_count, uint64_t* total_mem) {
  if (!c) return fail(S2M_ERR_INVALID_ARG, "ctx is NULL");
  if (name && name_len) { strncpy(name, c->prop.name, name_len - 1); name[name_len - 1] = 0; }
  if (sm_count) *sm_count = c->prop.multiProcessorCount;
  if (total_mem) *total_mem = c->prop.totalGlobalMem;
  return S2M_OK;
}

// ------------------------------------------------------------------ module
constexpr bool kK1PackedDefault = true;  // see s2m_pvec.h; S2M_K1_PACKED=0/1 overrides
constexpr bool kK1PackedSqrtDefault = true;
constexpr int kK1PackedMaxTinySize = 64;  // expression nodes of one evaluation (torus.sdf3d: 21, p_key.sdf3d: 206)
constexpr int kK1PackedMinScore = 4;     // transcendental calls in the SDF (mandelmesh.frag: 7; the .sdf3d examples: 0)
constexpr unsigned kK1RowsDefault = 2;  // measured: mandelbulb K1 -0.6 %, torus K1 -7 %; +10-20 % NVRTC time

struct s2m_module {
  std::string cuda_source, log;
  static constexpr int kMaxParts = 3;   // kernels_jit.cuh S2M_JIT_PART: K1 | K4a | diagnostic kernels
  int n_parts = 0;                      // 3, or 1 when everything was compiled as one program (S2M_JIT_SPLIT=0)
  std::vector<char> cubin[kMaxParts];
  CUmodule_t mod[kMaxParts] = {nullptr, nullptr, nullptr};
  CUfunction_t k1 = nullptr, k4 = nullptr, k_eval = nullptr, k_probe = nullptr, k_eval2 = nullptr;
  s2m_ctx* ctx = nullptr;
  unsigned k1_rows = 1;  // grid rows per K1 thread (S2M_K1_ROWS the kernels were compiled with)
  bool k1_packed = false;  // K1 evaluates corner pairs in f32x2 arithmetic (s2m_pvec.h)
  double ms_frontend = 0, ms_nvrtc = 0, ms_load = 0;
};

namespace {

void unload_parts(s2m_module* m) {
  if (!m->ctx) return;
  for (CUmodule_t& mod : m->mod)
    if (mod) { cudaSetDevice(m->ctx->device); driver().cuModuleUnload(mod); mod = nullptr; }
  m->k1 = m->k4 = m->k_eval = m->k_probe = m->k_eval2 = nullptr;
}

// cubins -> CUmodules and kernel handles on m->ctx's device
int load_parts(s2m_module* m) {
  using namespace s2m_internal;
  const double t0 = now_ms();
  CUDA_TRY(cudaSetDevice(m->ctx->device));
  CUresult_t cr;
  for (int k = 0; k < m->n_parts; ++k) {
    cr = driver().cuModuleLoadData(&m->mod[k], m->cubin[k].data());
    if (cr) { unload_parts(m); return fail(S2M_ERR_CUDA, "cuModuleLoadData: " + cu_err(cr)); }
  }
  const bool split = m->n_parts > 1;
  struct { CUfunction_t* f; const char* n; int part; bool wanted; } fns[] = {
      {&m->k1, "s2m_k1_slab", 0, true}, {&m->k4, "s2m_k4_vertices", 1, true}, {&m->k_eval, "s2m_k_eval", 2, true},
      {&m->k_probe, "s2m_k_cost_probe", 2, true}, {&m->k_eval2, "s2m_k_eval2", 2, m->k1_packed}};
  for (auto& f : fns) {
    if (!f.wanted) continue;
    cr = driver().cuModuleGetFunction(f.f, m->mod[split ? f.part : 0], f.n);
    if (cr) { unload_parts(m); return fail(S2M_ERR_CUDA, std::string("cuModuleGetFunction(") + f.n + "): " + cu_err(cr)); }
  }
  m->ms_load = now_ms() - t0;
  return S2M_OK;
}

// How K1 evaluates the SDF.  `packed_text` is the front-end's packed (f32x2)
// translation, empty for "one corner per evaluation".
struct K1Plan {
  std::string packed_text;
  bool packed_sqrt = kK1PackedSqrtDefault;  // the refinement step of sqrt in f32x2 as well (S2M_K1_PACKED=2)
  bool heavy = false;                       // >= kK1PackedMinScore transcendental calls
  bool packed() const { return !packed_text.empty(); }
};

// Default (measured on B200, tools/k1_ab.py): packed when the SDF is dominated by transcendental
// functions (mandelbulb K1 41.4 -> 40.3 ms: the polynomials halve, but the kernel then waits on
// latency with 48-62 registers) or when it is tiny (torus K1 13.5 -> 12.3 ms); not for mid-sized
// primitive compositions, whose packed kernels need 75-160 registers (martin_cube 24 -> 41 ms).
// The front-end reports both measures in the first comment line of the packed text.
// S2M_K1_PACKED=0 / 1 / 2 overrides: scalar / pairs / pairs + packed sqrt.
K1Plan plan_k1(std::string packed_text) {
  K1Plan plan;
  int score = 0, size = 0;
  const size_t at = packed_text.find("// s2m-packed-score: ");
  if (at != std::string::npos) sscanf(packed_text.c_str() + at + 21, "%d %d", &score, &size);
  plan.heavy = score >= kK1PackedMinScore;
  bool use = kK1PackedDefault && (plan.heavy || size <= kK1PackedMaxTinySize);
  if (const char* e = getenv("S2M_K1_PACKED")) {
    use = atoi(e) != 0;
    if (use) plan.packed_sqrt = atoi(e) >= 2;
  }
  if (use) plan.packed_text = std::move(packed_text);
  return plan;
}

// One NVRTC program: translation unit + options -> cubin (from S2M_CACHE_DIR when it is there).  Touches
// nothing but its arguments, so several parts compile concurrently.  Returns S2M_OK or S2M_ERR_NVRTC with
// the compiler's message in *error (*log holds the NVRTC log either way).
int compile_part(const std::string& source, const std::vector<std::string>& opts, std::vector<char>* cubin,
                 std::string* log, std::string* error) {
  using namespace s2m_internal;
  std::vector<const char*> opt_ptrs;
  for (const std::string& o : opts) opt_ptrs.push_back(o.c_str());
  const char* hdr_src[] = {kSrcMathH, kSrcVecH, kSrcSdfLibH, kSrcPvecH, kSrcScanCuh, kSrcKernelsJit};
  const char* hdr_name[] = {"s2m_math.h", "s2m_vec.h", "s2m_sdf3d_lib.h", "s2m_pvec.h", "s2m_scan.cuh", "kernels_jit.cuh"};

  // Optional on-disk cubin cache (S2M_CACHE_DIR): keyed by everything that determines the cubin -- the
  // generated translation unit, the embedded headers, the options (which name the part) and the NVRTC
  // version.  A serving process that sees the same SDF again skips the compile.
  std::string cache_path;
  cubin->clear();
  log->clear();
  if (const char* dir = getenv("S2M_CACHE_DIR")) {
    if (*dir) {
      int nv_major = 0, nv_minor = 0;
      nvrtcVersion(&nv_major, &nv_minor);
      std::string key = source;
      for (const char* h : hdr_src) { key += '\0'; key += h; }
      for (const std::string& o : opts) { key += '\0'; key += o; }
      key.push_back('\0');  // (a "\0..." literal would end at its first byte)
      key += "nvrtc " + std::to_string(nv_major) + "." + std::to_string(nv_minor) + " " + s2m_version();
      unsigned long long h1 = 1469598103934665603ull, h2 = 0x9e3779b97f4a7c15ull;  // two independent 64-bit FNV-1a style hashes
      for (unsigned char ch : key) { h1 = (h1 ^ ch) * 1099511628211ull; h2 = (h2 + ch) * 0xff51afd7ed558ccdull; h2 ^= h2 >> 29; }
      char name[64];
      snprintf(name, sizeof name, "/%016llx%016llx.cubin", h1, h2);
      cache_path = std::string(dir) + name;
      if (FILE* f = fopen(cache_path.c_str(), "rb")) {
        fseek(f, 0, SEEK_END);
        const long n = ftell(f);
        fseek(f, 0, SEEK_SET);
        if (n > 64) {
          cubin->resize((size_t)n);
          if (fread(cubin->data(), 1, (size_t)n, f) != (size_t)n || memcmp(cubin->data(), "\x7f" "ELF", 4) != 0) cubin->clear();
        }
        fclose(f);
      }
    }
  }
  if (!cubin->empty()) {
    *log = "cubin loaded from " + cache_path;
    return S2M_OK;
  }

  nvrtcProgram prog = nullptr;
  nvrtcResult r = nvrtcCreateProgram(&prog, source.c_str(), "sdf_module.cu", 6, hdr_src, hdr_name);
  if (r != NVRTC_SUCCESS) {
    *error = std::string("nvrtcCreateProgram: ") + nvrtcGetErrorString(r);
    return S2M_ERR_NVRTC;
  }
  r = nvrtcCompileProgram(prog, (int)opt_ptrs.size(), opt_ptrs.data());
  size_t ls = 0;
  nvrtcGetProgramLogSize(prog, &ls);
  if (ls > 1) { log->resize(ls); nvrtcGetProgramLog(prog, &(*log)[0]); log->resize(strlen(log->c_str())); }
  if (r != NVRTC_SUCCESS) {
    *error = std::string("NVRTC: ") + nvrtcGetErrorString(r) + "\n" + *log;
    nvrtcDestroyProgram(&prog);
    return S2M_ERR_NVRTC;
  }
  size_t cs = 0;
  nvrtcGetCUBINSize(prog, &cs);
  cubin->resize(cs);
  nvrtcGetCUBIN(prog, cubin->data());
  nvrtcDestroyProgram(&prog);
  if (!cache_path.empty()) {  // best effort: write to a temporary name, then rename (atomic on POSIX)
    const std::string tmp = cache_path + ".tmp" + std::to_string((long long)getpid()) + "." + std::to_string((unsigned long long)(uintptr_t)cubin);
    if (FILE* f = fopen(tmp.c_str(), "wb")) {
      const bool ok = fwrite(cubin->data(), 1, cubin->size(), f) == cubin->size();
      fclose(f);
      if (!ok || rename(tmp.c_str(), cache_path.c_str()) != 0) remove(tmp.c_str());
    }
  }
  return S2M_OK;
}

// Generated text + plan -> the module's cubins.  The translation unit is compiled as three programs
// that differ in -DS2M_JIT_PART (kernels_jit.cuh: K1 / K4a / diagnostic kernels), concurrently on three
// host threads: each carries only the copies of the user's SDF its kernels inline, and NVRTC compiles
// separate programs in parallel (8-vCPU build container, median of 5: martin_cube.sdf3d 1170 -> 670 ms,
// mandelmesh.frag 1050 -> 810 ms, p_key 560 -> 520 ms, torus unchanged at 550 ms; K4a's part is the longest).
// S2M_JIT_SPLIT=0 compiles one program with everything, as one would offline with nvcc.
int build_cubins(s2m_module* m, const std::string& user, const K1Plan& plan, uint32_t flags, std::string* error) {
  m->k1_packed = plan.packed();
  m->cuda_source = std::string("#include \"s2m_sdf3d_lib.h\"\n#include \"s2m_scan.cuh\"\n") +
                   "namespace s2m_user {\nusing namespace s2m;\n" + user + "\n}  // namespace s2m_user\n";
  if (plan.packed()) {
    m->cuda_source += std::string(plan.packed_sqrt ? "#define S2M_PACKED_SQRT 1\n" : "") +
                      "#include \"s2m_pvec.h\"\n#define S2M_K1_PACKED 1\nnamespace s2m_user_p {\nusing namespace s2m;\n" + plan.packed_text +
                      "\n}  // namespace s2m_user_p\n";
    if (getenv("S2M_TEST_BREAK_PACKED"))  // lets tests/test_capi.py exercise the fallback in s2m_module_compile
      m->cuda_source += "#error packed form rejected on request (S2M_TEST_BREAK_PACKED)\n";
  }
  m->cuda_source += "#include \"kernels_jit.cuh\"\n";

  std::vector<std::string> opts = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo",
                                   (flags & S2M_COMPILE_ALLOW_FMA) ? "--fmad=true" : "--fmad=false"};
  if (const char* e = getenv("S2M_K1_UNROLL"))  // experiment knob, see kernels_jit.cuh
    opts.push_back(std::string("-DS2M_K1_UNROLL=") + (atoi(e) == 1 ? "1" : "4"));
  // a packed evaluation carries twice the state: for a heavy SDF one row per thread and a 48-register
  // cap (5 resident blocks) measured best (mandelbulb: 62 registers / 4 blocks otherwise)
  const bool packed_heavy = plan.packed() && plan.heavy;
  m->k1_rows = packed_heavy ? 1u : kK1RowsDefault;
  if (const char* e = getenv("S2M_K1_ROWS")) m->k1_rows = atoi(e) == 2 ? 2u : 1u;  // experiment knob
  opts.push_back("-DS2M_K1_ROWS=" + std::to_string(m->k1_rows));
  if (const char* e = getenv("S2M_K1_MINBLOCKS"))  // experiment knob
    opts.push_back("-DS2M_K1_MINBLOCKS=" + std::to_string(std::max(1, std::min(8, atoi(e)))));
  else if (packed_heavy)
    opts.push_back("-DS2M_K1_MINBLOCKS=5");

  const char* split = getenv("S2M_JIT_SPLIT");
  const int n_parts = (split && atoi(split) == 0) ? 1 : s2m_module::kMaxParts;
  m->n_parts = n_parts;
  int status[s2m_module::kMaxParts] = {0, 0, 0};
  std::string logs[s2m_module::kMaxParts], errors[s2m_module::kMaxParts];
  auto work = [&](int k) {
    std::vector<std::string> o = opts;
    if (n_parts > 1) o.push_back("-DS2M_JIT_PART=" + std::to_string(k + 1));
    status[k] = compile_part(m->cuda_source, o, &m->cubin[k], &logs[k], &errors[k]);
  };
  std::vector<std::thread> threads;
  for (int k = 1; k < n_parts; ++k) {
    try { threads.emplace_back(work, k); }
    catch (const std::system_error&) { work(k); }   // the process may not create threads: compile this part here
  }
  work(0);
  for (std::thread& t : threads) t.join();
  m->log.clear();
  for (int k = 0; k < n_parts; ++k)
    if (!logs[k].empty() && m->log.find(logs[k]) == std::string::npos) m->log += (m->log.empty() ? "" : "\n") + logs[k];
  for (int k = 0; k < n_parts; ++k)
    if (status[k] != S2M_OK) {
      for (auto& c : m->cubin) c.clear();
      *error = errors[k];
      return status[k];
    }
  return S2M_OK;
}

}  // namespace

extern "C" int s2m_module_compile(s2m_ctx* ctx, const s2m_shader* shader, uint32_t flags, s2m_module** out) {
  using namespace s2m_internal;
  if (!shader || !out) return fail(S2M_ERR_INVALID_ARG, "s2m_module_compile: NULL argument");
  *out = nullptr;
  std::unique_ptr<s2m_module> m(new s2m_module());
  m->ctx = ctx;
  double t0 = now_ms();
  std::string user, user_packed, err;
  if (shader->kind == S2M_SRC_CUDA) {
    user = shader->source;
  } else {
    int st = s2m_frontend::lower_to_cuda(*shader, &user, &err, &user_packed);
    if (st != S2M_OK) return fail(st, err);
  }
  double t1 = now_ms();
  m->ms_frontend = t1 - t0;
  K1Plan plan = plan_k1(std::move(user_packed));
  std::string packed_log;
  int st = build_cubins(m.get(), user, plan, flags, &err);
  if (st != S2M_OK && plan.packed()) {  // NVRTC rejected the packed form: keep its diagnostics, compile the scalar kernels only
    packed_log = "packed (f32x2) form rejected, K1 falls back to one corner per evaluation:\n" + m->log + "\n";
    plan.packed_text.clear();
    st = build_cubins(m.get(), user, plan, flags, &err);
  }
  if (st != S2M_OK) return fail(st, err);
  if (!packed_log.empty()) m->log = packed_log + m->log;
  double t2 = now_ms();
  m->ms_nvrtc = t2 - t1;
  if (ctx) {
    st = load_parts(m.get());
    if (st != S2M_OK) return st;
  }
  *out = m.release();
  return S2M_OK;
}

extern "C" int s2m_module_instantiate(const s2m_module* compiled, s2m_ctx* ctx, s2m_module** out) {
  using namespace s2m_internal;
  if (!compiled || !ctx || !out) return fail(S2M_ERR_INVALID_ARG, "s2m_module_instantiate: NULL argument");
  *out = nullptr;
  if (compiled->n_parts < 1) return fail(S2M_ERR_STATE, "s2m_module_instantiate: the module holds no cubin");
  std::unique_ptr<s2m_module> m(new s2m_module());
  m->cuda_source = compiled->cuda_source;
  m->log = compiled->log;
  m->n_parts = compiled->n_parts;
  for (int k = 0; k < compiled->n_parts; ++k) m->cubin[k] = compiled->cubin[k];
  m->k1_rows = compiled->k1_rows;
  m->k1_packed = compiled->k1_packed;
  m->ms_frontend = compiled->ms_frontend;
  m->ms_nvrtc = compiled->ms_nvrtc;
  m->ctx = ctx;
  int st = load_parts(m.get());
  if (st != S2M_OK) return st;
  *out = m.release();
  return S2M_OK;
}

extern "C" const char* s2m_module_log(const s2m_module* m) { return m ? m->log.c_str() : ""; }
extern "C" const char* s2m_module_cuda_source(const s2m_module* m) { return m ? m->cuda_source.c_str() : ""; }
extern "C" int s2m_module_cubin(const s2m_module* m, const void** data, size_t* size) { return s2m_module_cubin_part(m, 0, data, size); }
extern "C" int s2m_module_cubin_part(const s2m_module* m, int part, const void** data, size_t* size) {
  if (!m || !data || !size) return fail(S2M_ERR_INVALID_ARG, "NULL argument");
  if (part < 0 || part >= m->n_parts) return fail(S2M_ERR_INVALID_ARG, "s2m_module_cubin_part: the module has " + std::to_string(m->n_parts) + " part(s)");
  *data = m->cubin[part].data(); *size = m->cubin[part].size();
  return S2M_OK;
}
extern "C" double s2m_module_compile_ms(const s2m_module* m, int which) {
  if (!m) return 0;
  return which == 0 ? m->ms_frontend : (which == 1 ? m->ms_nvrtc : m->ms_load);
}
extern "C" void s2m_module_free(s2m_module* m) {
  if (!m) return;
  unload_parts(m);
  delete m;
}

// ------------------------------------------------------------------ grid bookkeeping
namespace {
struct GridDev {  // must match S2mGrid in kernels_jit.cuh
  float bmin[3];
  float size[3];
  float eps;
  unsigned res[3];
  unsigned pitch_x;
  unsigned rows;
  unsigned long long plane_stride;
};
struct SlabViewDev {  // must match S2mSlabView
  const float* slab;
  unsigned first_plane;
  unsigned n_planes;
};
// K1 block shape (bx, by), bx*by = 256; S2M_K1_BLOCK=8x32 overrides the default for experiments
void k1_block_shape(unsigned* bx, unsigned* by) {
  struct Shape { unsigned x = 8, y = 32; };
  static const Shape shape = [] {  // initialised once, thread-safe (contexts may be driven from several host threads)
    Shape sh;
    if (const char* e = getenv("S2M_K1_BLOCK")) {
      unsigned a = 0, b = 0;
      if (sscanf(e, "%ux%u", &a, &b) == 2 && a && b && a * b <= 256 && (a * b) % 32 == 0 && a % 8 == 0) { sh.x = a; sh.y = b; }
    }
    return sh;
  }();
  *bx = shape.x; *by = shape.y;
}
struct VertexOutDev {  // must match S2mVertexOut
  float* pos; float* nrm; unsigned long long* key; unsigned char* nibble; unsigned* cand_vrank;
  unsigned long long* status; unsigned* ticket; unsigned long long* n_vertices; unsigned long long* n_halo;
};

int make_grid(const s2m_mesh_params* p, GridDev* g) {
  if (!p) return fail(S2M_ERR_INVALID_ARG, "params is NULL");
  if (p->struct_size != sizeof(s2m_mesh_params)) return fail(S2M_ERR_INVALID_ARG, "s2m_mesh_params.struct_size mismatch");
  for (int a = 0; a < 3; ++a) {
    if (p->dims[a] < 2 || p->dims[a] > 65535u)
      return fail(S2M_ERR_INVALID_ARG, "dims must be in [2, 65535] (cell coordinates are u16 in the reference key, mesh.rs:214)");
    g->bmin[a] = p->bb_min[a];
    volatile float v = (float)(p->dims[a] - 1u);      // dualcontour.wgsl:23
    volatile float extent = p->bb_max[a] - p->bb_min[a];
    g->size[a] = extent / v;                          // :24
    g->res[a] = p->dims[a];
  }
  g->eps = p->eps;
  g->pitch_x = (p->dims[0] + 1u + 31u) & ~31u;
  g->rows = p->dims[1] + 1u;
  g->plane_stride = (unsigned long long)g->pitch_x * g->rows;
  return S2M_OK;
}

int launch(CUfunction_t f, dim3 grid, dim3 block, cudaStream_t st, void** args, const char* name, unsigned dyn_smem = 0) {
  CUresult_t r = driver().cuLaunchKernel(f, grid.x, grid.y, grid.z, block.x, block.y, block.z, dyn_smem, st, args, nullptr);
  if (r) return fail(S2M_ERR_CUDA, std::string("cuLaunchKernel(") + name + "): " + cu_err(r));
  return S2M_OK;
}
}  // namespace

// ------------------------------------------------------------------ result
struct s2m_result {
  s2m_ctx* ctx = nullptr;
  s2m_module* mod = nullptr;
  s2m_mesh_params params{};
  GridDev grid{};
  uint32_t z_first = 0, nz = 0, label_add = 0, halo = 0, words_x = 0;
  uint64_t n_cand = 0, n_vert_total = 0, n_halo = 0, n_quads = 0, n_invalid = 0;
  float* h_pos = nullptr; float* h_nrm = nullptr; uint64_t* h_key = nullptr; uint8_t* h_nib = nullptr;
  uint64_t* h_quads = nullptr; uint64_t* h_cand = nullptr; uint64_t* h_invalid = nullptr; float* h_halo_pos = nullptr;
  int64_t global_base = 0;
  uint64_t n_invalid_records = 0;
  uint64_t cap_v = 0, cap_q = 0;   // capacity (elements) of the pinned vertex / quad blocks
  bool streamed = false;           // vertex (and quad) chunks were copied while later chunks computed
  bool quads_done = false;         // K4b ran inside begin (single-slab s2m_mesh_run)
  s2m_timings t{};
  double wall0 = 0;
  size_t ev_used = 0;              // events taken from the ctx pool by this run
  struct Span { int kind; size_t e0, e1; };  // kind: 0 K1, 1 K2, 2 K3, 3 K4a, 4 K4b, 5 copy
  std::vector<Span> spans;
  bool finished = false;
  // two-call form: the last chunk's vertex copy is issued by finish(), after the caller's count exchange --
  // a bulk copy in flight delays every small device->host read behind it (PCIe), including that exchange's result
  bool deferred_copy = false;
  uint64_t deferred_v0 = 0;
  bool quads_u32() const { return (params.flags & S2M_MESH_QUADS_U32) != 0; }
  size_t quad_bytes() const { return quads_u32() ? 16 : 32; }  // bytes per quad, device and host
};

extern "C" void s2m_result_free(s2m_result* r) {
  if (!r) return;
  if (r->ctx) {
    for (void* p : {(void*)r->h_pos, (void*)r->h_nrm, (void*)r->h_key, (void*)r->h_nib, (void*)r->h_quads, (void*)r->h_cand, (void*)r->h_invalid, (void*)r->h_halo_pos})
      if (p) r->ctx->release_pinned(p);
    if (!r->finished) r->ctx->busy = false;
  }
  delete r;
}

extern "C" int s2m_params_from_cli(uint32_t resolution, float bounds, s2m_mesh_params* out, int* rounded) {
  if (!out) return fail(S2M_ERR_INVALID_ARG, "out is NULL");
  memset(out, 0, sizeof *out);
  out->struct_size = sizeof *out;
  uint32_t res = resolution ? resolution : 256u;  // main.rs:141
  int r = 0;
  if (__builtin_popcount(res) > 1) {             // main.rs:142-148: 2 << ilog2(res)
    res = 2u << (31 - __builtin_clz(res));
    r = 1;
  }
  if (rounded) *rounded = r;
  float b = (bounds > 0.0f) ? bounds : 2.0f;     // main.rs:150 unwrap_or(2.0)
  volatile float v = b * 0.5f;                   // lib.rs:115-118
  for (int a = 0; a < 3; ++a) { out->bb_min[a] = 0.0f - v; out->bb_max[a] = 0.0f + v; out->dims[a] = res; }
  out->eps = 0.0001f;                            // main.rs:165
  return S2M_OK;
}

namespace {

// One z-chunk of the pipeline.  Slices are local to the slab: slice 0 == true slice r->z_first.
struct Chunk { uint32_t z0, nzc; };

size_t take_event(s2m_ctx* c, s2m_result* r) {
  if (r->ev_used == c->ev_pool.size()) {
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    c->ev_pool.push_back(e);
  }
  return r->ev_used++;
}
#define SPAN_BEGIN(kind_, stream_) \
  const size_t span_e0__ = take_event(c, r); CUDA_TRY(cudaEventRecord(c->ev_pool[span_e0__], stream_)); const int span_kind__ = kind_
#define SPAN_END(stream_) \
  do { const size_t e1__ = take_event(c, r); CUDA_TRY(cudaEventRecord(c->ev_pool[e1__], stream_)); r->spans.push_back({span_kind__, span_e0__, e1__}); } while (0)

int ensure_pinned_outputs(s2m_ctx* c, s2m_result* r, uint64_t need_v, uint64_t need_q, bool want_quads) {
  if (need_v > r->cap_v || !r->h_pos) {
    for (void* p : {(void*)r->h_pos, (void*)r->h_nrm, (void*)r->h_key, (void*)r->h_nib}) if (p) c->release_pinned(p);
    r->h_pos = (float*)c->lease_pinned(need_v * 12);
    r->h_nrm = (float*)c->lease_pinned(need_v * 12);
    r->h_key = (uint64_t*)c->lease_pinned(need_v * 8);
    r->h_nib = (uint8_t*)c->lease_pinned(need_v);
    if (!r->h_pos || !r->h_nrm || !r->h_key || !r->h_nib) return fail(S2M_ERR_OOM, "cudaHostAlloc for vertex output failed");
    r->cap_v = need_v;
  }
  if (want_quads && (need_q > r->cap_q || !r->h_quads)) {
    if (r->h_quads) c->release_pinned(r->h_quads);
    r->h_quads = (uint64_t*)c->lease_pinned(need_q * r->quad_bytes());
    if (!r->h_quads) return fail(S2M_ERR_OOM, "cudaHostAlloc for quad output failed");
    r->cap_q = need_q;
  }
  return S2M_OK;
}

// device -> pinned host copies of vertices [v0, v1) (indices include the halo) and quads [q0, q1)
int copy_out(s2m_ctx* c, s2m_result* r, cudaStream_t st, uint64_t v0, uint64_t v1, uint64_t q0, uint64_t q1) {
  v0 = std::max<uint64_t>(v0, r->n_halo);
  if (v1 > v0) {
    const uint64_t n = v1 - v0, h = v0 - r->n_halo;
    CUDA_TRY(cudaMemcpyAsync(r->h_pos + 3 * h, c->v_pos.as<float>() + 3 * v0, n * 12, cudaMemcpyDeviceToHost, st));
    if (!(r->params.flags & S2M_MESH_NO_NORMALS))
      CUDA_TRY(cudaMemcpyAsync(r->h_nrm + 3 * h, c->v_nrm.as<float>() + 3 * v0, n * 12, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(r->h_key + h, c->v_key.as<unsigned long long>() + v0, n * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(r->h_nib + h, c->v_nib.as<unsigned char>() + v0, n, cudaMemcpyDeviceToHost, st));
  }
  if (q1 > q0) {
    const size_t qb = r->quad_bytes();
    CUDA_TRY(cudaMemcpyAsync(reinterpret_cast<char*>(r->h_quads) + qb * q0, c->quads.as<char>() + qb * q0, (q1 - q0) * qb, cudaMemcpyDeviceToHost, st));
  }
  return S2M_OK;
}

int launch_k4b(s2m_ctx* c, s2m_result* r, cudaStream_t s, uint64_t v_begin, uint64_t v_end, uint64_t quad_base, long long index_offset) {
  unsigned long long* d_cnt = c->counters.as<unsigned long long>();
  v_begin = std::max<uint64_t>(v_begin, r->n_halo);
  if (v_end <= v_begin) return S2M_OK;
  const unsigned tiles = s2m_k4b_tiles(v_end - v_begin);
  int st;
  const size_t qb = r->quad_bytes();
  if ((st = c->quads.ensure_preserve((size_t)(quad_base + (v_end - v_begin) * 3) * qb + 64, (size_t)quad_base * qb, s))) return st;
  if ((st = c->scratch.ensure(((size_t)tiles + 8) * 8 + 64))) return st;
  CUDA_TRY(cudaMemsetAsync(c->scratch.p, 0, ((size_t)tiles + 8) * 8 + 64, s));
  S2mK4bArgs a{};
  a.vert_key = c->v_key.as<unsigned long long>(); a.vert_nibble = c->v_nib.as<unsigned char>();
  a.v_begin = v_begin; a.v_end = v_end; a.quad_base = quad_base;
  a.cand_mask = c->cand_mask.as<uint32_t>(); a.word_prefix = c->word_prefix.as<uint32_t>(); a.cand_vrank = c->cand_vrank.as<uint32_t>();
  a.words_x = r->words_x; a.res_y = r->grid.res[1]; a.z_first = r->z_first; a.label_add = r->label_add;
  a.index_offset = index_offset;
  a.quads = c->quads.as<unsigned long long>(); a.quads32 = r->quads_u32() ? c->quads.as<unsigned>() : nullptr;
  a.status = c->scratch.as<unsigned long long>() + 1;
  a.ticket = reinterpret_cast<unsigned*>(c->scratch.p); a.n_quads = d_cnt + C_NQUAD; a.n_invalid = d_cnt + C_NINVALID;
  if (r->params.flags & S2M_MESH_KEEP_INVALID) {
    if ((st = c->invalid.ensure(kInvalidCapacity * 48))) return st;
    a.invalid_records = c->invalid.as<unsigned long long>(); a.invalid_cursor = d_cnt + C_INVALID_CURSOR; a.invalid_capacity = kInvalidCapacity;
  }
  SPAN_BEGIN(4, s);
  int e = s2m_launch_k4b(&a, s);
  if (e) return fail(S2M_ERR_CUDA, std::string("k4_quads launch: ") + cudaGetErrorString((cudaError_t)e));
  SPAN_END(s);
  r->t.launches += 1;
  return S2M_OK;
}

int read_counters(s2m_ctx* c, cudaStream_t s) {
  // not a cudaMemcpy: see k_publish in kernels_static.cu
  int e = s2m_launch_publish(c->counters.as<unsigned long long>(), c->h_counters, C_COUNT, s);
  if (e) return fail(S2M_ERR_CUDA, std::string("k_publish launch: ") + cudaGetErrorString((cudaError_t)e));
  ++c->publish_launches;
  CUDA_TRY(cudaStreamSynchronize(s));
  return S2M_OK;
}

void finalize_timings(s2m_ctx* c, s2m_result* r) {
  float acc[6] = {0, 0, 0, 0, 0, 0};
  for (const auto& sp : r->spans) {
    float ms = 0;
    cudaEventElapsedTime(&ms, c->ev_pool[sp.e0], c->ev_pool[sp.e1]);
    acc[sp.kind] += ms;
  }
  r->t.k1_slab_ms = acc[0]; r->t.k2_classify_ms = acc[1]; r->t.k3_compact_ms = acc[2];
  r->t.k4_vertices_ms = acc[3]; r->t.k4_quads_ms = acc[4]; r->t.d2h_ms = acc[5];
  if (getenv("S2M_TRACE")) {  // device timeline: when each span started / ended relative to the first launch
    static const char* names[6] = {"K1", "K2", "K3", "K4a", "K4b", "copy"};
    for (const auto& sp : r->spans) {
      float t0 = 0, t1 = 0;
      cudaEventElapsedTime(&t0, c->ev[0], c->ev_pool[sp.e0]);
      cudaEventElapsedTime(&t1, c->ev[0], c->ev_pool[sp.e1]);
      fprintf(stderr, "[s2m timeline] %-4s %9.3f -> %9.3f ms (%7.3f)\n", names[sp.kind], t0, t1, t1 - t0);
    }
  }
  float ms = 0;
  cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]); r->t.device_ms = ms;
  cudaEventElapsedTime(&ms, c->ev[0], c->ev[2]); r->t.total_ms = ms;
}

// The pipeline.  For every z-chunk: K1 slab -> K2 classify -> [count] -> K3 compact -> K4a vertices
// -> [count] -> (K4b quads -> [count]) -> async copy of the chunk's vertices (and quads) into pinned
// host memory on the copy stream, overlapping the next chunk's kernels.  [count] = the host reads
// 8 bytes to size the next launch.
int mesh_begin_impl(s2m_ctx* c, s2m_module* m, const s2m_mesh_params* p, bool fuse_quads, s2m_result** out) {
  if (!c || !m || !p || !out) return fail(S2M_ERR_INVALID_ARG, "s2m_mesh_begin: NULL argument");
  *out = nullptr;
  if (!m->k1 || m->ctx != c) return fail(S2M_ERR_STATE, "module was not compiled for this ctx");
  if (c->busy) return fail(S2M_ERR_STATE, "a previous s2m_mesh_begin on this ctx has not been finished or freed");
  CUDA_TRY(cudaSetDevice(c->device));
  std::unique_ptr<s2m_result> rp(new s2m_result());
  s2m_result* r = rp.get();
  r->ctx = c; r->mod = m; r->params = *p;
  int st = make_grid(p, &r->grid);
  if (st) return st;
  const GridDev& g = r->grid;
  const bool all = p->flags & S2M_MESH_ALL_SLICES;
  const bool dense = p->flags & S2M_MESH_EXACT_DENSE;
  const uint32_t zlast = all ? g.res[2] : g.res[2] - 1u;  // the last slice is never read back
  uint32_t zb = p->z_begin, ze = p->z_end;
  if (zb == 0 && ze == 0) ze = zlast;
  ze = std::min(ze, zlast);
  if (zb > ze) zb = ze;
  r->label_add = all ? 0u : 1u;
  r->halo = (zb > 0 && zb < ze) ? 1u : 0u;
  r->z_first = zb - r->halo;
  r->nz = ze - r->z_first;
  r->words_x = (g.res[0] + 31u) / 32u;
  if (r->halo) fuse_quads = false;  // a halo means there are lower slabs: the global vertex base comes later
  const float min_size = std::min(g.size[0], std::min(g.size[1], g.size[2]));
  // Candidate band.  The reference's corner coordinate (min + size) and the slab's (bmin + size*(i+1))
  // differ by at most 1 ulp of the coordinate; in voxels that is 2^-23 * max|coordinate| / size.  The
  // default band tolerates field changes of up to 256x that move (a true SDF changes by 1x), and is
  // never thinner than 1/16 voxel.
  float ulp_voxels = 0.0f;
  for (int a = 0; a < 3; ++a)
    ulp_voxels = std::max(ulp_voxels, std::max(std::fabs(p->bb_min[a]), std::fabs(p->bb_max[a])) * 1.1920929e-7f / g.size[a]);
  const float tau_default = std::max(0.0625f, 256.0f * 1.7320508f * ulp_voxels);
  const float tau = (p->tau_voxels > 0.0f ? p->tau_voxels : tau_default) * min_size;

  Trace tr;
  c->busy = true;
  c->publish_launches = 0;
  struct BusyGuard { s2m_ctx* c; bool keep = false; ~BusyGuard() { if (!keep) c->busy = false; } } guard{c};
  r->wall0 = now_ms();
  cudaStream_t s = c->stream;
  unsigned long long* d_cnt = c->counters.as<unsigned long long>();
  CUDA_TRY(cudaMemsetAsync(d_cnt, 0, C_COUNT * 8, s));
  CUDA_TRY(cudaEventRecord(c->ev[0], s));

  const unsigned long long words_per_slice = (unsigned long long)g.res[1] * r->words_x;
  const unsigned long long n_words = words_per_slice * r->nz;
  if (r->nz > 0) {
    if ((st = c->cand_mask.ensure((n_words + 16) * 4))) return st;
    if ((st = c->word_prefix.ensure((n_words + 16) * 4))) return st;
  }
  // ---- chunk plan: bounded slab, and enough chunks that the copies hide behind later chunks
  std::vector<Chunk> chunks;
  const unsigned long long plane_bytes = g.plane_stride * 4ull;
  if (r->nz > 0) {
    unsigned long long budget = p->slab_budget_bytes ? p->slab_budget_bytes : (4ull << 30);
    uint32_t zc = 0;
    for (;;) {
      const unsigned long long max_planes = std::max<unsigned long long>(2, budget / plane_bytes);
      zc = (uint32_t)std::min<unsigned long long>(r->nz, max_planes - 1);
      if (dense) break;
      st = c->slab.ensure((unsigned long long)(zc + 1) * plane_bytes);
      if (st == S2M_OK) break;
      if (st != S2M_ERR_OOM || zc <= 1) return st;
      budget = (unsigned long long)(zc + 1) * plane_bytes / 2;
    }
    uint32_t n_chunks = (r->nz + zc - 1) / zc;
    if (!dense && !getenv("S2M_NO_CHUNK_OVERLAP")) {
      // a slab that fits the budget in one piece is still cut into 2-4 chunks when it is large enough
      // (>= 0.15 G voxels per chunk) for the two-stream overlap and the early output copies to pay
      const double voxels = (double)g.res[0] * g.res[1] * r->nz;
      const uint32_t want = (uint32_t)std::min(4.0, voxels / 1.5e8);
      n_chunks = std::max(n_chunks, std::min(want, r->nz));
    }
    const uint32_t even = (r->nz + n_chunks - 1) / n_chunks;  // equal chunks instead of a short last one
    for (uint32_t z0 = 0; z0 < r->nz; z0 += even) chunks.push_back({z0, std::min(even, r->nz - z0)});
  }
  r->t.chunks = (uint32_t)chunks.size();
  // Two slab (and class-plane) buffers when there is more than one chunk: K1/K2 of chunk c+1 then run
  // on the producer stream while the consumer stream works through K3/K4a/K4b of chunk c.
  bool pipelined = !dense && chunks.size() > 1 && !getenv("S2M_NO_CHUNK_OVERLAP");
  if (pipelined) {
    const size_t max_planes = (size_t)chunks[0].nzc + 1;
    if (c->slab2.ensure(max_planes * plane_bytes) != S2M_OK) { cudaGetLastError(); pipelined = false; }  // not enough memory: one buffer, no overlap
  }
  // pinned output sized from the previous run on this ctx (if any): lets chunk copies start early
  const bool can_stream = c->hint_nv > 0;
  if (can_stream) {
    if ((st = ensure_pinned_outputs(c, r, c->hint_nv + c->hint_nv / 32 + 4096, c->hint_nq + c->hint_nq / 32 + 4096, fuse_quads))) return st;
    r->streamed = true;
  }
  tr.mark("setup");

  uint64_t cand_done = 0, vert_done = 0, quad_done = 0;
  std::vector<uint32_t> dense_row;
  cudaStream_t ps = pipelined ? c->prod_stream : s;   // producer stream (K1, K2)
  if (pipelined) CUDA_TRY(cudaStreamWaitEvent(ps, c->ev[EV_BEGIN], 0));  // after the counter reset
  const bool from_slab = p->flags & S2M_MESH_CLASSIFY_FROM_SLAB;
  const unsigned cls_words = g.pitch_x / 32u;
  // K1 + K2 of chunk ci into slab buffer ci % 2 (buffer 0 when not pipelined)
  auto produce = [&](size_t ci) -> int {
    const Chunk ch = chunks[ci];
    const int buf = pipelined ? (int)(ci & 1) : 0;
    DevBuf& slab_buf = buf ? c->slab2 : c->slab;
    DevBuf& cls_buf = buf ? c->cls2 : c->cls;
    int st2;
    GridDev gd = g;
    float* slab = slab_buf.as<float>();
    unsigned first_plane = r->z_first + ch.z0, n_planes = ch.nzc + 1;
    unsigned cw = cls_words;
    void* cls = nullptr;
    if (!from_slab) {
      if ((st2 = cls_buf.ensure((size_t)n_planes * g.rows * cls_words * 8 + 64))) return st2;
      cls = cls_buf.p;
    }
    if (pipelined && ci >= 2) CUDA_TRY(cudaStreamWaitEvent(ps, c->ev[EV_CONSUMED0 + buf], 0));  // K4a of chunk ci-2 has read this buffer
    CUDA_TRY(cudaMemsetAsync(d_cnt + C_CHUNK_CAND0 + buf, 0, 8, ps));
    float tau_arg = tau;
    void* a1[] = {&gd, &slab, &first_plane, &n_planes, &tau_arg, &cls, &cw};
    unsigned bx, by;
    k1_block_shape(&bx, &by);
    dim3 grid1((g.pitch_x + 4u * bx - 1u) / (4u * bx), (g.rows + by * m->k1_rows - 1u) / (by * m->k1_rows), n_planes);
    {
      SPAN_BEGIN(0, ps);
      if ((st2 = launch(m->k1, grid1, dim3(bx, by, 1), ps, a1, "s2m_k1_slab"))) return st2;
      SPAN_END(ps);
    }
    S2mK2Args a2{};
    a2.slab = slab; a2.pitch_x = g.pitch_x; a2.plane_stride = g.plane_stride;
    a2.res_x = g.res[0]; a2.res_y = g.res[1]; a2.nz_chunk = ch.nzc; a2.tau = tau;
    a2.cand_mask = c->cand_mask.as<uint32_t>() + words_per_slice * ch.z0; a2.words_x = r->words_x; a2.total = d_cnt + C_CHUNK_CAND0 + buf;
    a2.cls = cls; a2.cls_words = cls_words;
    {
      SPAN_BEGIN(1, ps);
      int e2 = from_slab ? s2m_launch_k2(&a2, ps) : s2m_launch_k2_bits(&a2, ps);
      if (e2) return fail(S2M_ERR_CUDA, std::string("k2_classify launch: ") + cudaGetErrorString((cudaError_t)e2));
      SPAN_END(ps);
    }
    if (pipelined) CUDA_TRY(cudaEventRecord(c->ev[EV_PRODUCED0 + buf], ps));
    r->t.launches += 2;
    return S2M_OK;
  };
  if (!dense && !chunks.empty() && (st = produce(0))) return st;
  for (size_t ci = 0; ci < chunks.size(); ++ci) {
    const Chunk ch = chunks[ci];
    const int buf = pipelined ? (int)(ci & 1) : 0;
    const unsigned long long chunk_words = words_per_slice * ch.nzc;
    uint32_t* mask_chunk = c->cand_mask.as<uint32_t>() + words_per_slice * ch.z0;
    unsigned slab_first_plane = 0, slab_n_planes = 0;
    uint64_t cand_total = cand_done;
    if (!dense) {
      // the next chunk's K1/K2 are queued before this chunk's counts are waited for
      if (pipelined && ci + 1 < chunks.size() && (st = produce(ci + 1))) return st;
      if (pipelined) CUDA_TRY(cudaStreamWaitEvent(s, c->ev[EV_PRODUCED0 + buf], 0));
      slab_first_plane = r->z_first + ch.z0; slab_n_planes = ch.nzc + 1;
      // ---- [count] candidates of this chunk
      if ((st = read_counters(c, s))) return st;
      cand_total = cand_done + c->h_counters[C_CHUNK_CAND0 + buf];
      if (!pipelined && ci + 1 < chunks.size()) { /* single buffer: the next K1 is issued after this chunk's K4a (below) */ }
    } else {
      // reference-cost mode: every cell of the chunk is a candidate
      if (dense_row.empty()) {
        dense_row.assign(r->words_x, 0xffffffffu);
        if (g.res[0] % 32u) dense_row.back() = (1u << (g.res[0] % 32u)) - 1u;
      }
      std::vector<uint32_t> host(chunk_words);
      for (unsigned long long i = 0; i < chunk_words; i += r->words_x) memcpy(&host[i], dense_row.data(), r->words_x * 4);
      CUDA_TRY(cudaMemcpyAsync(mask_chunk, host.data(), chunk_words * 4, cudaMemcpyHostToDevice, s));
      CUDA_TRY(cudaStreamSynchronize(s));
      cand_total = cand_done + (unsigned long long)g.res[0] * g.res[1] * ch.nzc;
    }
    const uint64_t n_cand = cand_total - cand_done;
    if (cand_total >= 0xffffffffull) return fail(S2M_ERR_UNSUPPORTED, "more than 2^32-1 candidate cells in one slab; split it with z_begin/z_end");
    const unsigned k3_tiles = s2m_k3_tiles(chunk_words);
    const unsigned k4_tiles = (unsigned)((n_cand + 127) / 128);
    const size_t status_words = (size_t)k3_tiles + k4_tiles + 16;
    if ((st = c->status.ensure(status_words * 8))) return st;
    if ((st = c->cand_key.ensure_preserve((cand_total + 1) * 8, cand_done * 8, s))) return st;
    if ((st = c->cand_vrank.ensure_preserve((cand_total + 1) * 4, cand_done * 4, s))) return st;
    if ((st = c->v_pos.ensure_preserve((vert_done + n_cand + 1) * 12, vert_done * 12, s))) return st;
    if ((st = c->v_nrm.ensure_preserve((vert_done + n_cand + 1) * 12, vert_done * 12, s))) return st;
    if ((st = c->v_key.ensure_preserve((vert_done + n_cand + 1) * 8, vert_done * 8, s))) return st;
    if ((st = c->v_nib.ensure_preserve(vert_done + n_cand + 16, vert_done, s))) return st;
    CUDA_TRY(cudaMemsetAsync(c->status.p, 0, status_words * 8, s));
    unsigned long long* status = c->status.as<unsigned long long>();
    unsigned* tickets = reinterpret_cast<unsigned*>(status);  // words 0..1: tickets; 2..: tile status
    // ---- K3
    {
      S2mK3Args a3{};
      a3.cand_mask = mask_chunk; a3.n_words = chunk_words; a3.words_x = r->words_x; a3.res_y = g.res[1];
      a3.z_offset = r->z_first + ch.z0; a3.word_prefix = c->word_prefix.as<uint32_t>() + words_per_slice * ch.z0;
      a3.cand_key = c->cand_key.as<unsigned long long>(); a3.base = cand_done;
      a3.status = status + 2; a3.ticket = tickets;
      SPAN_BEGIN(2, s);
      int e3 = s2m_launch_k3(&a3, s);
      if (e3) return fail(S2M_ERR_CUDA, std::string("k3_compact launch: ") + cudaGetErrorString((cudaError_t)e3));
      SPAN_END(s);
      r->t.launches += 1;
    }
    // ---- K4a
    if (k4_tiles) {
      GridDev gd = g;
      const unsigned long long* ck = c->cand_key.as<unsigned long long>() + cand_done;
      unsigned long long nc = n_cand, vbase = vert_done;
      unsigned label_add = r->label_add, halo_below = r->halo ? (r->z_first + 1u) : 0u;
      unsigned want_normals = ((p->flags & S2M_MESH_NO_NORMALS) ? 0u : 1u) | ((p->flags & S2M_MESH_CONSISTENT_CORNERS) ? 2u : 0u);  // K4a's mode bits
      SlabViewDev sv{(buf ? c->slab2 : c->slab).as<float>(), slab_first_plane, slab_n_planes};
      VertexOutDev vo{c->v_pos.as<float>(), c->v_nrm.as<float>(), c->v_key.as<unsigned long long>(), c->v_nib.as<unsigned char>(),
                      c->cand_vrank.as<unsigned>() + cand_done, status + 2 + k3_tiles, tickets + 1, d_cnt + C_NVERT, d_cnt + C_NHALO};
      void* a4[] = {&gd, &ck, &nc, &vbase, &label_add, &halo_below, &want_normals, &sv, &vo};
      SPAN_BEGIN(3, s);
      // (capping K4a's blocks per SM with unused dynamic shared memory, to keep K1 blocks of the next chunk
      // resident beside them, was measured: K4a 4.3 -> 6.4 ms and the run got 2 ms slower)
      if ((st = launch(m->k4, dim3(k4_tiles), dim3(128), s, a4, "s2m_k4_vertices"))) return st;
      SPAN_END(s);
      r->t.launches += 1;
    }
    if (pipelined) CUDA_TRY(cudaEventRecord(c->ev[EV_CONSUMED0 + buf], s));  // slab buffer `buf` may be overwritten
    else if (!dense && ci + 1 < chunks.size() && (st = produce(ci + 1))) return st;
    // ---- [count] vertices so far
    uint64_t vert_total = vert_done;
    if (k4_tiles) {
      if ((st = read_counters(c, s))) return st;
      vert_total = c->h_counters[C_NVERT];
      r->n_halo = c->h_counters[C_NHALO];
    }
    // ---- K4b (single-slab runs only: the global vertex base is 0)
    uint64_t quad_total = quad_done;
    if (fuse_quads && vert_total > vert_done) {
      if ((st = launch_k4b(c, r, s, vert_done, vert_total, quad_done, 0))) return st;
      if ((st = read_counters(c, s))) return st;
      quad_total = c->h_counters[C_NQUAD];
    }
    // ---- copy this chunk's output while the next chunk computes
    if (r->streamed) {
      const uint64_t own = vert_total - std::min<uint64_t>(vert_total, r->n_halo);
      if (own > r->cap_v || (fuse_quads && quad_total > r->cap_q)) {
        r->streamed = false;  // the hint was too small: everything is copied at the end instead
      } else if (!fuse_quads && ci + 1 == chunks.size()) {
        r->deferred_copy = true;  // see s2m_result::deferred_copy
        r->deferred_v0 = vert_done;
      } else {
        const size_t e = take_event(c, r);
        CUDA_TRY(cudaEventRecord(c->ev_pool[e], s));
        CUDA_TRY(cudaStreamWaitEvent(c->copy_stream, c->ev_pool[e], 0));
        SPAN_BEGIN(5, c->copy_stream);
        if ((st = copy_out(c, r, c->copy_stream, vert_done, vert_total, quad_done, fuse_quads ? quad_total : quad_done))) return st;
        SPAN_END(c->copy_stream);
      }
    }
    cand_done = cand_total; vert_done = vert_total; quad_done = quad_total;
    tr.mark("chunk done");
  }
  r->n_cand = cand_done;
  r->n_vert_total = vert_done;
  const uint64_t n_own = r->n_vert_total - r->n_halo;
  if (fuse_quads) {
    r->quads_done = true;
    r->n_quads = quad_done;
    if ((st = read_counters(c, s))) return st;
    r->n_invalid = c->h_counters[C_NINVALID];
  }
  CUDA_TRY(cudaEventRecord(c->ev[1], s));  // device_ms: every kernel of begin() has finished
  // ---- whatever has not been streamed is copied now
  if (!r->streamed) {
    if ((st = ensure_pinned_outputs(c, r, n_own, fuse_quads ? r->n_quads : 0, fuse_quads))) return st;
    SPAN_BEGIN(5, s);
    if ((st = copy_out(c, r, s, 0, r->n_vert_total, 0, fuse_quads ? r->n_quads : 0))) return st;
    SPAN_END(s);
  }
  if (p->flags & S2M_MESH_NO_NORMALS) memset(r->h_nrm, 0, n_own * 12);
  if (p->flags & S2M_MESH_KEEP_CANDIDATES) {
    r->h_cand = (uint64_t*)c->lease_pinned(r->n_cand * 8);
    if (!r->h_cand) return fail(S2M_ERR_OOM, "cudaHostAlloc for candidate list failed");
    if (r->n_cand) CUDA_TRY(cudaMemcpyAsync(r->h_cand, c->cand_key.p, r->n_cand * 8, cudaMemcpyDeviceToHost, s));
  }
  if (r->n_halo) {  // the halo slice's positions, so that this slab's triangles can be written without the slab below
    r->h_halo_pos = (float*)c->lease_pinned(r->n_halo * 12);
    if (!r->h_halo_pos) return fail(S2M_ERR_OOM, "cudaHostAlloc for halo positions failed");
    CUDA_TRY(cudaMemcpyAsync(r->h_halo_pos, c->v_pos.p, r->n_halo * 12, cudaMemcpyDeviceToHost, s));
  }
  c->hint_nv = n_own;
  tr.mark("begin done");
  guard.keep = true;
  *out = rp.release();
  return S2M_OK;
}

}  // namespace

extern "C" int s2m_mesh_begin(s2m_ctx* c, s2m_module* m, const s2m_mesh_params* p, s2m_result** out) {
  return mesh_begin_impl(c, m, p, false, out);
}

extern "C" int s2m_mesh_finish(s2m_result* r, int64_t global_vertex_base) {
  if (!r || !r->ctx) return fail(S2M_ERR_INVALID_ARG, "s2m_mesh_finish: NULL result");
  if (r->finished) return fail(S2M_ERR_STATE, "s2m_mesh_finish called twice");
  s2m_ctx* c = r->ctx;
  Trace tr;
  CUDA_TRY(cudaSetDevice(c->device));
  cudaStream_t s = c->stream;
  int st;
  r->global_base = global_vertex_base;
  if (r->quads_u32() && (global_vertex_base < 0 || (uint64_t)global_vertex_base + (r->n_vert_total - r->n_halo) > 0xffffffffull))
    return fail(S2M_ERR_UNSUPPORTED, "S2M_MESH_QUADS_U32: vertex indices of this slab do not fit 32 bits");
  if (r->deferred_copy && r->streamed) {
    SPAN_BEGIN(5, c->copy_stream);
    if ((st = copy_out(c, r, c->copy_stream, r->deferred_v0, r->n_vert_total, 0, 0))) return st;
    SPAN_END(c->copy_stream);
    r->deferred_copy = false;
  }
  if (!r->quads_done) {
    // ---- K4b over all own vertices with the global base, then the quad copy
    if ((st = launch_k4b(c, r, s, 0, r->n_vert_total, 0, (long long)global_vertex_base - (long long)r->n_halo))) return st;
    CUDA_TRY(cudaEventRecord(c->ev[1], s));
    if ((st = read_counters(c, s))) return st;
    r->n_quads = c->h_counters[C_NQUAD];
    r->n_invalid = c->h_counters[C_NINVALID];
    if ((st = ensure_pinned_outputs(c, r, r->cap_v, r->n_quads, true))) return st;
    SPAN_BEGIN(5, s);
    if ((st = copy_out(c, r, s, 0, 0, 0, r->n_quads))) return st;
    SPAN_END(s);
    r->quads_done = true;
  } else if (global_vertex_base != 0) {
    return fail(S2M_ERR_STATE, "quads were already emitted with base 0 (s2m_mesh_run); use s2m_mesh_begin for multi-slab runs");
  }
  c->hint_nq = r->n_quads;
  if (r->params.flags & S2M_MESH_KEEP_INVALID) {
    if ((st = read_counters(c, s))) return st;
    r->n_invalid_records = std::min<uint64_t>(c->h_counters[C_INVALID_CURSOR], kInvalidCapacity);
    r->h_invalid = (uint64_t*)c->lease_pinned(r->n_invalid_records * 48 + 48);
    if (!r->h_invalid) return fail(S2M_ERR_OOM, "cudaHostAlloc for the invalid-quad list failed");
    if (r->n_invalid_records) {
      CUDA_TRY(cudaMemcpyAsync(r->h_invalid, c->invalid.p, r->n_invalid_records * 48, cudaMemcpyDeviceToHost, s));
      CUDA_TRY(cudaStreamSynchronize(s));
      // the device appends in completion order; the reference reports them in (vertex, edge) order
      struct Rec { uint64_t v[6]; };
      Rec* recs = reinterpret_cast<Rec*>(r->h_invalid);
      std::sort(recs, recs + r->n_invalid_records, [](const Rec& x, const Rec& y) { return x.v[0] != y.v[0] ? x.v[0] < y.v[0] : x.v[1] < y.v[1]; });
    }
  }
  {  // everything (both streams) done -> total span
    const size_t e = take_event(c, r);
    CUDA_TRY(cudaEventRecord(c->ev_pool[e], c->copy_stream));
    CUDA_TRY(cudaStreamWaitEvent(s, c->ev_pool[e], 0));
  }
  CUDA_TRY(cudaEventRecord(c->ev[2], s));
  CUDA_TRY(cudaStreamSynchronize(s));
  CUDA_TRY(cudaStreamSynchronize(c->copy_stream));
  tr.mark("finish: copies done");
  r->t.host_wall_ms = now_ms() - r->wall0;
  r->t.launches += c->publish_launches;
  c->publish_launches = 0;
  finalize_timings(c, r);
  r->finished = true;
  c->busy = false;
  return S2M_OK;
}

extern "C" int s2m_mesh_run(s2m_ctx* c, s2m_module* m, const s2m_mesh_params* p, s2m_result** out) {
  int st = mesh_begin_impl(c, m, p, true, out);
  if (st) return st;
  st = s2m_mesh_finish(*out, 0);
  if (st) { s2m_result_free(*out); *out = nullptr; }
  return st;
}

extern "C" int s2m_result_get(const s2m_result* r, s2m_result_info* o) {
  if (!r || !o) return fail(S2M_ERR_INVALID_ARG, "NULL argument");
  memset(o, 0, sizeof *o);
  o->n_vertices = r->n_vert_total - r->n_halo;
  o->n_halo_vertices = r->n_halo;
  o->n_quads = r->n_quads;
  o->n_invalid_quads = r->n_invalid;
  o->n_candidates = r->n_cand;
  o->positions = r->h_pos; o->normals = r->h_nrm; o->cell_keys = r->h_key; o->sign_nibbles = r->h_nib;
  o->quads = r->quads_u32() ? nullptr : r->h_quads; o->quads32 = r->quads_u32() ? reinterpret_cast<const uint32_t*>(r->h_quads) : nullptr;
  o->candidates = r->h_cand;
  o->invalid_records = r->h_invalid; o->n_invalid_records = r->n_invalid_records;
  o->halo_positions = r->h_halo_pos; o->global_vertex_base = r->global_base;
  o->timings = r->t;
  return S2M_OK;
}

// ------------------------------------------------------------------ diagnostics
extern "C" int s2m_eval_points(s2m_ctx* c, s2m_module* m, const float* xyz, uint64_t n, float* out) {
  if (!c || !m || !xyz || !out) return fail(S2M_ERR_INVALID_ARG, "NULL argument");
  if (!m->k1 || m->ctx != c) return fail(S2M_ERR_STATE, "module was not compiled for this ctx");
  if (n == 0) return S2M_OK;
  CUDA_TRY(cudaSetDevice(c->device));
  DevBuf in, o;
  int st;
  if ((st = in.ensure(n * 12))) return st;
  if ((st = o.ensure(n * 4))) { in.release(); return st; }
  cudaStream_t s = c->stream;
  cudaError_t e = cudaMemcpyAsync(in.p, xyz, n * 12, cudaMemcpyHostToDevice, s);
  const float* pin = in.as<float>(); float* pout = o.as<float>(); unsigned long long nn = n;
  void* args[] = {&pin, &pout, &nn};
  if (e == cudaSuccess) st = launch(m->k_eval, dim3((unsigned)((n + 255) / 256)), dim3(256), s, args, "s2m_k_eval");
  if (e == cudaSuccess && !st) e = cudaMemcpyAsync(out, o.p, n * 4, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess && !st) e = cudaStreamSynchronize(s);
  in.release(); o.release();
  if (st) return st;
  if (e != cudaSuccess) return fail(S2M_ERR_CUDA, std::string("s2m_eval_points: ") + cudaGetErrorString(e));
  return S2M_OK;
}

extern "C" int s2m_module_is_packed(const s2m_module* m) { return m && m->k1_packed ? 1 : 0; }

extern "C" int s2m_eval_pairs(s2m_ctx* c, s2m_module* m, const float* xyz_a, const float* xyz_b, uint64_t n, float* out_a, float* out_b,
                              uint8_t* disagreed) {
  if (!c || !m || !xyz_a || !xyz_b || !out_a || !out_b || !disagreed) return fail(S2M_ERR_INVALID_ARG, "NULL argument");
  if (!m->k1 || m->ctx != c) return fail(S2M_ERR_STATE, "module was not compiled for this ctx");
  if (!m->k_eval2) return fail(S2M_ERR_UNSUPPORTED, "module has no packed (f32x2) form");
  if (n == 0) return S2M_OK;
  CUDA_TRY(cudaSetDevice(c->device));
  DevBuf in, o;  // in: a | b;  o: out_a | out_b | dv
  int st;
  if ((st = in.ensure(n * 24))) return st;
  if ((st = o.ensure(n * 9))) { in.release(); return st; }
  cudaStream_t s = c->stream;
  cudaError_t e = cudaMemcpyAsync(in.p, xyz_a, n * 12, cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(in.as<char>() + n * 12, xyz_b, n * 12, cudaMemcpyHostToDevice, s);
  const float* pa = in.as<float>(); const float* pb = pa + n * 3;
  float* oa = o.as<float>(); float* ob = oa + n; unsigned char* dv = reinterpret_cast<unsigned char*>(ob + n);
  unsigned long long nn = n;
  void* args[] = {&pa, &pb, &oa, &ob, &dv, &nn};
  if (e == cudaSuccess) st = launch(m->k_eval2, dim3((unsigned)((n + 127) / 128)), dim3(128), s, args, "s2m_k_eval2");
  if (e == cudaSuccess && !st) e = cudaMemcpyAsync(out_a, oa, n * 4, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess && !st) e = cudaMemcpyAsync(out_b, ob, n * 4, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess && !st) e = cudaMemcpyAsync(disagreed, dv, n, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess && !st) e = cudaStreamSynchronize(s);
  in.release(); o.release();
  if (st) return st;
  if (e != cudaSuccess) return fail(S2M_ERR_CUDA, std::string("s2m_eval_pairs: ") + cudaGetErrorString(e));
  return S2M_OK;
}

extern "C" int s2m_debug_slab_plane(s2m_ctx* c, s2m_module* m, const s2m_mesh_params* p, uint32_t plane, float* out) {
  if (!c || !m || !p || !out) return fail(S2M_ERR_INVALID_ARG, "NULL argument");
  if (!m->k1 || m->ctx != c) return fail(S2M_ERR_STATE, "module was not compiled for this ctx");
  if (c->busy) return fail(S2M_ERR_STATE, "ctx is busy");
  GridDev g;
  int st = make_grid(p, &g);
  if (st) return st;
  if (plane > g.res[2]) return fail(S2M_ERR_INVALID_ARG, "plane out of range");
  CUDA_TRY(cudaSetDevice(c->device));
  if ((st = c->slab.ensure(g.plane_stride * 4))) return st;
  float* slab = c->slab.as<float>();
  unsigned first_plane = plane, n_planes = 1, cls_words = 0;
  float tau_arg = 0.0f;
  void* cls = nullptr;
  void* a1[] = {&g, &slab, &first_plane, &n_planes, &tau_arg, &cls, &cls_words};
  unsigned bx, by;
  k1_block_shape(&bx, &by);
  if ((st = launch(m->k1, dim3((g.pitch_x + 4u * bx - 1u) / (4u * bx), (g.rows + by * m->k1_rows - 1u) / (by * m->k1_rows), 1), dim3(bx, by, 1), c->stream, a1, "s2m_k1_slab"))) return st;
  CUDA_TRY(cudaMemcpy2DAsync(out, (size_t)(g.res[0] + 1) * 4, slab, (size_t)g.pitch_x * 4, (size_t)(g.res[0] + 1) * 4, g.rows,
                             cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return S2M_OK;
}

extern "C" int s2m_read_device_words(s2m_ctx* c, const void* device_words, uint32_t n, uint64_t* out, void* cuda_stream) {
  if (!c || !device_words || !out || n == 0 || n > 32) return fail(S2M_ERR_INVALID_ARG, "s2m_read_device_words: bad argument (n must be 1..32)");
  CUDA_TRY(cudaSetDevice(c->device));
  cudaStream_t s = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : c->stream;
  int e = s2m_launch_publish(static_cast<const unsigned long long*>(device_words), c->h_words, n, s);
  if (e) return fail(S2M_ERR_CUDA, std::string("k_publish launch: ") + cudaGetErrorString((cudaError_t)e));
  CUDA_TRY(cudaStreamSynchronize(s));
  for (uint32_t i = 0; i < n; ++i) out[i] = c->h_words[i];
  return S2M_OK;
}

extern "C" int s2m_cost_probe(s2m_ctx* c, s2m_module* m, const s2m_mesh_params* p, uint32_t planes, double* cost_out) {
  if (!c || !m || !p || !cost_out || planes == 0) return fail(S2M_ERR_INVALID_ARG, "bad argument");
  if (!m->k1 || m->ctx != c) return fail(S2M_ERR_STATE, "module was not compiled for this ctx");
  GridDev g;
  int st = make_grid(p, &g);
  if (st) return st;
  CUDA_TRY(cudaSetDevice(c->device));
  DevBuf buf;
  if ((st = buf.ensure((size_t)planes * 8 + 64))) return st;
  cudaStream_t s = c->stream;
  cudaError_t e = cudaMemsetAsync(buf.p, 0, (size_t)planes * 8 + 64, s);
  unsigned probe = 128, pl = planes;
  unsigned long long* cyc = buf.as<unsigned long long>();
  float* sink = reinterpret_cast<float*>(cyc + planes);
  void* args[] = {&g, &probe, &pl, &cyc, &sink};
  if (e == cudaSuccess) st = launch(m->k_probe, dim3(planes), dim3(256), s, args, "s2m_k_cost_probe");
  std::vector<unsigned long long> h(planes);
  if (e == cudaSuccess && !st) e = cudaMemcpyAsync(h.data(), cyc, (size_t)planes * 8, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess && !st) e = cudaStreamSynchronize(s);
  buf.release();
  if (st) return st;
  if (e != cudaSuccess) return fail(S2M_ERR_CUDA, std::string("s2m_cost_probe: ") + cudaGetErrorString(e));
  for (uint32_t i = 0; i < planes; ++i) cost_out[i] = (double)h[i];
  return S2M_OK;
}
