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
#include <atomic>
#include <chrono>
#include <cmath>
#include <sys/mman.h>
#include <sys/stat.h>
#include <cerrno>
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
    size_t want = bytes + (bytes >> 2) + 256;   // 25 % slack: a z-slab whose boundaries are refined between runs grows without a new allocation
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

// A growable pinned host array for one of a result's big outputs (positions, normals, keys, nibbles, quads).
// The virtual address range is reserved once (mmap, MAP_NORESERVE: costs nothing until touched) for the largest
// size the array can reach, and page-locked piece by piece (cudaHostRegister) as the run learns how much it
// needs: the array stays contiguous, never moves, and a FIRST run on a context can start copying chunk 0's
// vertices while later chunks are still being computed, without knowing the final size (cudaHostAlloc of the
// whole output up front is what a cold run used to wait for: ~0.3 ms per MB).  Regions go back to the
// context's pool with their pages still locked, so steady-state runs register nothing.
struct PinnedRegion {
  char* base = nullptr;
  size_t reserved = 0, registered = 0;
  bool used = false;
  int role = 0;   // which output array it serves (s2m_ctx::lease_region)
  std::vector<std::pair<size_t, size_t>> segs;  // registered (offset, length) pieces, for cudaHostUnregister
  int ensure(size_t bytes, size_t ahead) {
    if (bytes <= registered) return S2M_OK;
    if (bytes > reserved) return fail(S2M_ERR_OOM, "output exceeds the reserved host address range (" + std::to_string(reserved) + " B)");
    constexpr size_t kGrain = 2u << 20;
    size_t upto = std::min(reserved, (bytes + ahead + kGrain - 1) / kGrain * kGrain);
    cudaError_t e = cudaHostRegister(base + registered, upto - registered, cudaHostRegisterPortable);
    if (e != cudaSuccess && upto > bytes) {   // retry without the read-ahead
      cudaGetLastError();
      upto = std::min(reserved, (bytes + kGrain - 1) / kGrain * kGrain);
      e = cudaHostRegister(base + registered, upto - registered, cudaHostRegisterPortable);
    }
    if (e != cudaSuccess) { cudaGetLastError(); return fail(S2M_ERR_OOM, std::string("cudaHostRegister(") + std::to_string(upto - registered) + " B): " + cudaGetErrorString(e)); }
    segs.emplace_back(registered, upto - registered);
    registered = upto;
    return S2M_OK;
  }
  // device -> this array at byte offset `off`.  One cudaMemcpyAsync may not straddle two cudaHostRegister ranges
  // ("invalid argument"): the copy is cut at the boundaries of the registered pieces.
  int copy_from_device(size_t off, const void* src, size_t bytes, cudaStream_t st) {
    if (off + bytes > registered) return fail(S2M_ERR_STATE, "copy past the page-locked part of an output array");
    for (const auto& sg : segs) {
      const size_t a = std::max(off, sg.first), b = std::min(off + bytes, sg.first + sg.second);
      if (a >= b) continue;
      CUDA_TRY(cudaMemcpyAsync(base + a, static_cast<const char*>(src) + (a - off), b - a, cudaMemcpyDeviceToHost, st));
    }
    return S2M_OK;
  }
  void destroy() {
    for (auto& sg : segs) cudaHostUnregister(base + sg.first);
    segs.clear();
    if (base) munmap(base, reserved);
    base = nullptr; reserved = registered = 0;
  }
};

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
  DevBuf slab, cls, slab2, cls2, coords, cand_mask, seg_count, word_prefix, cand_key, cand_vrank, status, counters;
  DevBuf v_pos, v_nrm, v_key, v_nib, quads, scratch, invalid;
  std::vector<PinnedBlock> pinned;            // small outputs (candidate list, invalid records, halo positions)
  std::vector<std::unique_ptr<PinnedRegion>> regions;  // big outputs
  unsigned long long* h_counters = nullptr;  // pinned + mapped, 16 words: [buf] = candidate count of the chunk in slab buffer buf (written by K2)
  unsigned long long* h_words = nullptr;     // pinned + mapped, 32 words (s2m_read_device_words)
  unsigned long long* h_slots = nullptr;     // pinned + mapped, 8 words per z-chunk: written by K4b's last block
  size_t h_slots_cap = 0;                    // chunks
  size_t va_limit = ~(size_t)0;              // largest host address range mmap has been seen to refuse, halved (lease_region)
  DevBuf chunk_tot;                          // u64 [n_chunks + 1][2]: {vertices, quads} emitted up to the end of each z-chunk
  cudaEvent_t ev[16]{};
  std::vector<cudaEvent_t> ev_pool;   // per-launch timing events, grown on demand
  bool busy = false;  // a begin() without finish()/free() is outstanding

  // A region whose address range can hold `reserve` bytes.  `role` names the array (0 positions, 1 normals, 2 keys,
  // 3 nibbles, 4 quads): a run takes back the region the same array used last time, whose pages are already locked for
  // that array's size -- picking "any region that is large enough" let the arrays swap regions from run to run and
  // re-register tens of MB inside every run.
  PinnedRegion* lease_region(int role, size_t reserve) {
    PinnedRegion* best = nullptr;
    for (auto& g : regions)
      if (!g->used && g->role == role && g->reserved >= std::min(reserve, va_limit) && (!best || g->registered > best->registered)) best = g.get();
    if (best) { best->used = true; return best; }
    std::unique_ptr<PinnedRegion> g(new PinnedRegion());
    size_t want = std::max<size_t>(reserve, 64u << 20);
    want = (want + (2u << 20) - 1) / (2u << 20) * (2u << 20);
    void* m = MAP_FAILED;
    for (;; want /= 2) {   // an address-space limit (ulimit -v, strict overcommit): settle for less; an output that outgrows it is an OOM error
      want = std::min(want, va_limit);
      m = mmap(nullptr, want, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
      if (m != MAP_FAILED) break;
      if (want <= (64u << 20)) return nullptr;
      va_limit = want / 2;
    }
#ifdef MADV_HUGEPAGE
    madvise(m, want, MADV_HUGEPAGE);   // 2 MiB pages where the kernel grants them: fewer pages to lock and to map for DMA
#endif
    g->base = static_cast<char*>(m); g->reserved = want; g->used = true; g->role = role;
    regions.push_back(std::move(g));
    return regions.back().get();
  }
  int ensure_slots(size_t chunks) {
    if (chunks <= h_slots_cap) return S2M_OK;
    if (h_slots) cudaFreeHost(h_slots);
    h_slots = nullptr; h_slots_cap = 0;
    const size_t want = std::max<size_t>(chunks + chunks / 2, 64);
    if (cudaHostAlloc(reinterpret_cast<void**>(&h_slots), want * 64, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) {
      cudaGetLastError();
      return fail(S2M_ERR_OOM, "cudaHostAlloc for the chunk result slots failed");
    }
    h_slots_cap = want;
    return S2M_OK;
  }

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

// device counters (u64 words).  C_CHUNK_CANDb / C_K2_DONEb: candidates of the chunk K2 classified into slab buffer b and
// K2's block-completion counter for it (adjacent: one 16-byte memset resets both).
enum Counter { C_NHALO = 2, C_NINVALID = 4, C_INVALID_CURSOR = 8, C_CHUNK_CAND0 = 9, C_K2_DONE0 = 10, C_CHUNK_CAND1 = 11, C_K2_DONE1 = 12, C_COUNT = 16 };
enum CtxEvent { EV_BEGIN = 0, EV_KERNELS_DONE = 1, EV_ALL_DONE = 2, EV_PRODUCED0 = 4, EV_PRODUCED1 = 5, EV_CONSUMED0 = 6, EV_CONSUMED1 = 7, EV_K1DONE0 = 10, EV_K1DONE1 = 11 };
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
  for (DevBuf* b : {&c->slab, &c->cls, &c->slab2, &c->cls2, &c->cand_mask, &c->seg_count, &c->word_prefix, &c->cand_key, &c->cand_vrank, &c->status, &c->counters,
                    &c->v_pos, &c->v_nrm, &c->v_key, &c->v_nib, &c->quads, &c->scratch, &c->invalid})
    b->release();
  for (auto& b : c->pinned) cudaFreeHost(b.p);
  for (auto& g : c->regions) g->destroy();
  c->chunk_tot.release();
  if (c->h_counters) cudaFreeHost(c->h_counters);
  if (c->h_words) cudaFreeHost(c->h_words);
  if (c->h_slots) cudaFreeHost(c->h_slots);
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
// Slab-free default (S2M_MESH_NO_SLAB): an SDF of at most this many expression nodes and no transcendental-heavy body.
// Writing 4 B per corner costs K1 more than K4a saves by reading 5.5 of a candidate's 8 corners back (torus 2048^3:
// K1 is 43 % of HBM write bandwidth with the slab and issue-bound without it).  torus.sdf3d: 21 nodes, p_key.sdf3d: 206.
constexpr int kSlabFreeMaxSize = 256;
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
  unsigned k1_zpt = 1;   // planes per K1 thread (S2M_K1_ZPT)
  bool k1_coords = false;  // K1 reads its corner coordinates from the run's table (S2M_K1_COORDS)
  bool k1_packed = false;  // K1 evaluates corner pairs in f32x2 arithmetic (s2m_pvec.h)
  bool slab_free_default = false;  // cheap SDF: K1 writes no f32 slab, K4a evaluates all 8 corners (S2M_MESH_NO_SLAB is the default for this module)
  double ms_frontend = 0, ms_nvrtc = 0, ms_load = 0;
  uint64_t uid = next_uid();   // s2m_module_uid; s2m_module_instantiate copies the number of the compiled module
  static uint64_t next_uid() { static std::atomic<uint64_t> n{0}; return ++n; }
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
  bool slab_free = false;                   // cheap SDF: meshing defaults to the slab-free form (kSlabFreeMaxSize)
  std::string packed_text;
  bool packed_sqrt = kK1PackedSqrtDefault;  // the refinement step of sqrt in f32x2 as well (S2M_K1_PACKED=2)
  bool heavy = false;                       // >= kK1PackedMinScore transcendental calls
  int size = 0;                             // expression nodes of one evaluation (0: unknown, e.g. S2M_SRC_CUDA input)
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
  plan.size = size;
  plan.slab_free = at != std::string::npos && !plan.heavy && size > 0 && size <= kSlabFreeMaxSize;
  bool use = kK1PackedDefault && (plan.heavy || size <= kK1PackedMaxTinySize);
  if (const char* e = getenv("S2M_K1_PACKED")) {
    use = atoi(e) != 0;
    if (use) plan.packed_sqrt = atoi(e) >= 2;
  }
  if (use) plan.packed_text = std::move(packed_text);
  return plan;
}

// Where compiled cubins are kept between processes: $S2M_CACHE_DIR if set (set but empty: no cache), else
// $XDG_CACHE_HOME/sdf2mesh_b200, else $HOME/.cache/sdf2mesh_b200 (created on first use; no home directory: no cache).
// A one-shot CLI run of an SDF it has seen before then skips NVRTC (0.5-0.8 s) for a file read of ~1 ms.
std::string cubin_cache_dir() {
  if (const char* e = getenv("S2M_CACHE_DIR")) return e;
  std::string base;
  if (const char* x = getenv("XDG_CACHE_HOME")) base = x;
  if (base.empty()) {
    const char* h = getenv("HOME");
    if (!h || !*h) return "";
    base = std::string(h) + "/.cache";
    mkdir(base.c_str(), 0700);
  }
  const std::string dir = base + "/sdf2mesh_b200";
  if (mkdir(dir.c_str(), 0700) != 0 && errno != EEXIST) return "";
  return dir;
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
  const std::string cache_dir = cubin_cache_dir();
  if (const char* dir = cache_dir.empty() ? nullptr : cache_dir.c_str()) {
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
  m->slab_free_default = plan.slab_free;
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
  // Launch shape by the size of the SDF (measured on B200, profiles/r02_k1_ab.jsonl; every variant gives the same bits):
  //  * heavy packed (>= 4 transcendental calls; mandelbulb): one row per thread, 16 planes marched per thread with the
  //    coordinates from the run's table, 64-register cap (4 resident blocks).  (5 blocks / 48 registers and one plane
  //    per thread won by 1-2 % before the round-2 instruction diet; after it: K1 35.5 ms at one plane, 34.3 at 2,
  //    33.2 at 4, 32.8 at 8, 32.6 at 16, 32.8 at 32.)
  //  * tiny (<= 64 expression nodes; torus): two rows per thread and 16 planes marched per thread -- index, coordinate
  //    and class overhead is a third of its instructions (K1 11.4 -> 8.7 ms at 2048^3)
  //  * in between (p_key 206 nodes, martin_cube 683): one row per thread with a 64-register cap (two rows need 111 / 240
  //    registers); up to 400 nodes 8 planes marched per thread (p_key K1 10.6 -> 6.6 ms at 1024^3), above that ONE inlined
  //    copy of the SDF per thread instead of four (martin_cube 512^3: 5.0 -> 2.6 ms)
  const bool packed_heavy = plan.packed() && plan.heavy;
  const bool tiny = plan.size > 0 && plan.size <= kK1PackedMaxTinySize;
  const bool mid = !plan.heavy && !tiny;
  m->k1_rows = (packed_heavy || mid) ? 1u : kK1RowsDefault;
  if (const char* e = getenv("S2M_K1_ROWS")) m->k1_rows = atoi(e) == 2 ? 2u : 1u;  // experiment knob
  opts.push_back("-DS2M_K1_ROWS=" + std::to_string(m->k1_rows));
  m->k1_zpt = (tiny || packed_heavy) ? 16u : ((mid && plan.size > 0 && plan.size <= 400) ? 8u : 1u);
  if (const char* e = getenv("S2M_K1_ZPT")) m->k1_zpt = (unsigned)std::max(1, std::min(64, atoi(e)));  // experiment knob
  opts.push_back("-DS2M_K1_ZPT=" + std::to_string(m->k1_zpt));
  // Corner coordinates from the run's table instead of i2f + mul + add per coordinate (kernels_jit.cuh): the mandelbulb's
  // K1 40.0 -> 38.5 ms at 2048^3, the others within noise (profiles/r02_k1_ab.jsonl); a thread that marches through
  // planes computes its x and y once anyway.
  m->k1_coords = m->k1_zpt == 1 || packed_heavy;
  if (const char* e = getenv("S2M_K1_COORDS")) m->k1_coords = atoi(e) != 0;  // experiment knob
  if (m->k1_coords) opts.push_back("-DS2M_K1_COORDS=1");
  if (const char* e = getenv("S2M_K1_UNROLL"))  // experiment knob, see kernels_jit.cuh
    opts.push_back(std::string("-DS2M_K1_UNROLL=") + (atoi(e) == 1 ? "1" : "4"));
  else if (mid && plan.size > 400)
    opts.push_back("-DS2M_K1_UNROLL=1");
  if (const char* e = getenv("S2M_K1_MINBLOCKS"))  // experiment knob
    opts.push_back("-DS2M_K1_MINBLOCKS=" + std::to_string(std::max(1, std::min(8, atoi(e)))));
  else if (packed_heavy || mid)
    opts.push_back("-DS2M_K1_MINBLOCKS=4");
  if (const char* e = getenv("S2M_K1_GUARD")) { if (atoi(e) != 0) opts.push_back("-DS2M_K1_GUARD=1"); }  // experiment knob
  // (no policy sets it any more: the heavy packed kernel kept the guard until it got its plane loop; see s2m_k1_eval4)

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
  m->k1_zpt = compiled->k1_zpt;
  m->k1_coords = compiled->k1_coords;
  m->uid = compiled->uid;
  m->k1_packed = compiled->k1_packed;
  m->slab_free_default = compiled->slab_free_default;
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
// The run's corner coordinate table for K1 (k_coords): x | y | z, each long enough for every thread of the K1 grid
// (threads past the grid's edge are inactive but still load).  Filled on `stream`, in front of the first K1.
static int prepare_coords(s2m_ctx* c, const s2m_module* m, const GridDev& g, unsigned bx, unsigned by, cudaStream_t stream, const float* out[3]) {
  out[0] = out[1] = out[2] = nullptr;
  if (!m->k1_coords) return S2M_OK;
  const unsigned gx = (g.pitch_x + 4u * bx - 1u) / (4u * bx), gy = (g.rows + by * m->k1_rows - 1u) / (by * m->k1_rows);
  const unsigned nx = gx * 4u * bx + 4u, ny = (gy * by * m->k1_rows + 2u + 3u) & ~3u, nz = g.res[2] + 1u;
  int st = c->coords.ensure((size_t)(nx + ny + nz) * 4 + 64);
  if (st) return st;
  float* tab = c->coords.as<float>();
  if (s2m_launch_coords(tab, nx, ny, nz, g.bmin, g.size, stream) != 0) return fail(S2M_ERR_CUDA, "k_coords launch failed");
  out[0] = tab; out[1] = tab + nx; out[2] = tab + nx + ny;
  return S2M_OK;
}
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
  // big outputs: growable pinned regions (PinnedRegion); small ones: blocks of the pinned pool
  PinnedRegion *g_pos = nullptr, *g_nrm = nullptr, *g_key = nullptr, *g_nib = nullptr, *g_quads = nullptr;
  uint64_t* h_cand = nullptr; uint64_t* h_invalid = nullptr; float* h_halo_pos = nullptr;
  int64_t global_base = 0;
  uint64_t n_invalid_records = 0;
  uint64_t copied_v = 0, copied_q = 0;   // vertices (halo included) / quads whose device->host copy has been issued
  s2m_timings t{};
  double wall0 = 0;
  size_t ev_used = 0;              // events taken from the ctx pool by this run
  struct Span { int kind; size_t e0, e1; };  // kind: 0 K1, 1 K2, 2 K3, 3 K4a, 4 K4b, 5 copy
  std::vector<Span> spans;
  bool want_spans = false;         // S2M_MESH_TIMINGS / S2M_TRACE
  bool finished = false;
  bool quads_u32() const { return (params.flags & S2M_MESH_QUADS_U32) != 0; }
  bool relative() const { return (params.flags & S2M_MESH_RELATIVE_QUADS) != 0; }
  size_t quad_bytes() const { return quads_u32() ? 16 : 32; }  // bytes per quad, device and host
  template <class T> T* host(PinnedRegion* g) const { return g ? reinterpret_cast<T*>(g->base) : nullptr; }
};

extern "C" void s2m_result_free(s2m_result* r) {
  if (!r) return;
  if (s2m_ctx* c = r->ctx) {
    if (!r->finished) {   // copies into the regions may still be in flight: wait before the memory is handed to the next run
      cudaSetDevice(c->device);
      cudaStreamSynchronize(c->copy_stream);
      cudaStreamSynchronize(c->stream);
      cudaStreamSynchronize(c->prod_stream);   // a run that failed midway may have queued the next chunk's K1
      c->busy = false;
    }
    for (PinnedRegion* g : {r->g_pos, r->g_nrm, r->g_key, r->g_nib, r->g_quads}) if (g) g->used = false;
    for (void* p : {(void*)r->h_cand, (void*)r->h_invalid, (void*)r->h_halo_pos}) if (p) c->release_pinned(p);
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
// per-kernel CUDA-event spans: only with S2M_MESH_TIMINGS (or S2M_TRACE).  Two event records around every launch keep
// consecutive kernels of a stream from overlapping head and tail, and reading ~70 event pairs back costs 0.35 ms per run.
#define SPAN_BEGIN(kind_, stream_) \
  size_t span_e0__ = 0; const int span_kind__ = kind_; \
  if (r->want_spans) { span_e0__ = take_event(c, r); CUDA_TRY(cudaEventRecord(c->ev_pool[span_e0__], stream_)); }
#define SPAN_END(stream_) \
  do { if (r->want_spans) { const size_t e1__ = take_event(c, r); CUDA_TRY(cudaEventRecord(c->ev_pool[e1__], stream_)); r->spans.push_back({span_kind__, span_e0__, e1__}); } } while (0)

// Device -> pinned host copies of what the chunks up to now have added: vertices [copied_v, vert_total) (local
// indices, the halo slice's first and never copied) and quads [copied_q, quad_total).  The pinned regions grow
// (cudaHostRegister) by what this copy needs plus as much again, so a first run registers a chunk ahead.
int copy_out(s2m_ctx* c, s2m_result* r, cudaStream_t st, uint64_t vert_total, uint64_t quad_total) {
  const uint64_t v0 = std::max<uint64_t>(r->copied_v, r->n_halo), v1 = vert_total;
  int e;
  if (v1 > v0) {
    const uint64_t n = v1 - v0, h = v0 - r->n_halo, own = v1 - r->n_halo;
    const bool normals = !(r->params.flags & S2M_MESH_NO_NORMALS);
    if ((e = r->g_pos->ensure(own * 12, n * 12))) return e;
    if (normals && (e = r->g_nrm->ensure(own * 12, n * 12))) return e;
    if ((e = r->g_key->ensure(own * 8, n * 8))) return e;
    if ((e = r->g_nib->ensure(own, n))) return e;
    if ((e = r->g_pos->copy_from_device(12 * h, c->v_pos.as<float>() + 3 * v0, n * 12, st))) return e;
    if (normals && (e = r->g_nrm->copy_from_device(12 * h, c->v_nrm.as<float>() + 3 * v0, n * 12, st))) return e;
    if ((e = r->g_key->copy_from_device(8 * h, c->v_key.as<unsigned long long>() + v0, n * 8, st))) return e;
    if ((e = r->g_nib->copy_from_device(h, c->v_nib.as<unsigned char>() + v0, n, st))) return e;
  }
  r->copied_v = std::max(r->copied_v, v1);
  if (quad_total > r->copied_q) {
    const size_t qb = r->quad_bytes();
    const uint64_t q0 = r->copied_q, n = quad_total - q0;
    if ((e = r->g_quads->ensure(quad_total * qb, n * qb))) return e;
    if ((e = r->g_quads->copy_from_device(qb * q0, c->quads.as<char>() + qb * q0, n * qb, st))) return e;
    r->copied_q = quad_total;
  }
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

// The pipeline.  For every z-chunk, on the PRODUCER stream: K1 (SDF once per corner -> f32 slab + 2-bit corner classes);
// on the CONSUMER stream: K2 (classes -> candidate bit per cell; its last block writes the chunk's candidate count into
// mapped host memory), K3 (compaction + rank table), K4a (the reference's per-cell arithmetic on the candidates ->
// vertices), K4b (quads, slab-relative indices; its last block writes the running totals into mapped host memory);
// on the COPY stream: the chunk's vertices and quads into pinned host memory.  The host waits for exactly two events
// per chunk -- "K2 done" (to size K4a's grid and the buffers) and "K4b done" (to size the copies) -- and both waits
// happen while the producer stream already runs K1 of the next chunk.  Ranges that depend on counts the host has not
// seen yet (vertices before / after this chunk, quads so far, halo vertices) are read by the kernels from device
// memory (c->chunk_tot, the counters).
int mesh_begin_impl(s2m_ctx* c, s2m_module* m, const s2m_mesh_params* p, s2m_result** out) {
  if (!c || !m || !p || !out) return fail(S2M_ERR_INVALID_ARG, "s2m_mesh_begin: NULL argument");
  *out = nullptr;
  if (!m->k1 || m->ctx != c) return fail(S2M_ERR_STATE, "module was not compiled for this ctx");
  if (c->busy) return fail(S2M_ERR_STATE, "a previous s2m_mesh_begin on this ctx has not been finished or freed");
  CUDA_TRY(cudaSetDevice(c->device));
  std::unique_ptr<s2m_result, void (*)(s2m_result*)> rp(new s2m_result(), [](s2m_result* x) { s2m_result_free(x); });
  s2m_result* r = rp.get();
  r->mod = m; r->params = *p;
  r->want_spans = (p->flags & S2M_MESH_TIMINGS) != 0 || getenv("S2M_TRACE") != nullptr;
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
  // Slab-free form: K1 writes only the corner classes, K4a evaluates all 8 corners of a candidate.  Chosen for SDFs
  // that are cheap to evaluate (the f32 slab costs more to write than K4a saves by reading it back); S2M_MESH_NO_SLAB
  // forces it, S2M_SLAB=0/1 in the environment overrides both (experiments).  K2 from the slab needs the slab.
  bool no_slab = !dense && ((p->flags & S2M_MESH_NO_SLAB) || m->slab_free_default);
  if (const char* e = getenv("S2M_SLAB")) no_slab = !dense && atoi(e) == 0;
  const bool from_slab = (p->flags & S2M_MESH_CLASSIFY_FROM_SLAB) != 0;
  if (from_slab) no_slab = false;

  Trace tr;
  r->ctx = c;
  c->busy = true;
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
    if ((st = c->seg_count.ensure(((unsigned long long)g.res[1] * r->nz * s2m_segs_x(r->words_x) + 16) * 4))) return st;
  }
  const unsigned long long segs_per_slice = (unsigned long long)g.res[1] * s2m_segs_x(r->words_x);
  // ---- chunk plan: bounded slab, and enough chunks that the copies hide behind later chunks
  std::vector<Chunk> chunks;
  const unsigned long long plane_bytes = g.plane_stride * 4ull;
  if (r->nz > 0) {
    // default 8 GiB per slab buffer (two of them when pipelined, of 180 GB): 2048^3 in 5 chunks, 25 launches; 4 GiB (9 chunks)
    // was 0.6 % slower, 3 GiB (11 chunks) 1.3 % (profiles/r02_k1_ab.jsonl)
    unsigned long long budget = p->slab_budget_bytes ? p->slab_budget_bytes : (8ull << 30);
    if (!p->slab_budget_bytes) if (const char* e = getenv("S2M_SLAB_BUDGET_GB")) budget = (unsigned long long)(std::max(0.01, atof(e)) * (double)(1ull << 30));  // experiment knob
    uint32_t zc = r->nz;
    if (!no_slab || p->slab_budget_bytes) {
      for (;;) {
        const unsigned long long max_planes = std::max<unsigned long long>(2, budget / plane_bytes);
        zc = (uint32_t)std::min<unsigned long long>(r->nz, max_planes - 1);
        if (dense || no_slab) break;
        st = c->slab.ensure((unsigned long long)(zc + 1) * plane_bytes);
        if (st == S2M_OK) break;
        if (st != S2M_ERR_OOM || zc <= 1) return st;
        budget = (unsigned long long)(zc + 1) * plane_bytes / 2;
      }
    }
    uint32_t n_chunks = (r->nz + zc - 1) / zc;
    if (!dense && !getenv("S2M_NO_CHUNK_OVERLAP")) {
      // a slab that fits the budget in one piece is still cut into chunks when it is large enough
      // (>= 0.15 G voxels per chunk) for the two-stream overlap and the early output copies to pay:
      // up to 4 with a slab, up to 8 without one (nothing is resident per chunk but the class planes)
      const double voxels = (double)g.res[0] * g.res[1] * r->nz;
      uint32_t want = (uint32_t)std::min(no_slab ? 8.0 : 4.0, voxels / 1.5e8);
      if (const char* e = getenv("S2M_CHUNKS_WANT")) want = (uint32_t)std::max(1, atoi(e));  // experiment knob
      n_chunks = std::max(n_chunks, std::min(want, r->nz));
    }
    const uint32_t even = (r->nz + n_chunks - 1) / n_chunks;  // equal chunks instead of a short last one
    for (uint32_t z0 = 0; z0 < r->nz; z0 += even) chunks.push_back({z0, std::min(even, r->nz - z0)});
    // Taper: what follows the LAST K1 launch -- K2, K3, K4a, K4b and the copy of the last chunk -- overlaps nothing and is
    // proportional to what that chunk holds, so the last chunk is cut into 1/2, 1/4, 1/8, 1/8 of its thickness; every
    // extra chunk costs five launches and two host waits.  It pays where the last chunk is full of surface: a slab that
    // ends inside the grid (N = 2: the lower slab of the mandelbulb ends in its densest slices, 2.2 ms of a 19.6 ms step
    // were that tail).  A slab that reaches the top of the grid usually ends in empty space, and there equal chunks are
    // as fast or faster (2048^3 mandelbulb 40.03 vs 40.12 ms, 1024^3 6.32 vs 6.61, torus 2048^3 9.74 vs 10.08, p_key
    // 1024^3 6.93 vs 7.08; profiles/r02_k1_ab.jsonl).  S2M_CHUNK_TAPER=0 / 1 forces it off / on.
    const bool taper = [&] {
      if (const char* e = getenv("S2M_CHUNK_TAPER")) return atoi(e) != 0;
      if (no_slab) return false;   // a cheap SDF's tail is short: torus 2048^3 at N = 2 5.14 ms with equal chunks, 5.70 tapered
      return r->z_first + r->nz < g.res[2] - ((p->flags & S2M_MESH_ALL_SLICES) ? 0u : 1u);
    }();
    if (taper && !dense && chunks.size() > 1 && chunks.back().nzc >= 32 && !getenv("S2M_NO_CHUNK_OVERLAP")) {
      const Chunk last = chunks.back();
      chunks.pop_back();
      uint32_t z0 = last.z0, left = last.nzc;
      for (int k = 0; k < 3 && left >= 16; ++k) {
        const uint32_t t = (left + 1) / 2;
        chunks.push_back({z0, t});
        z0 += t; left -= t;
      }
      if (left) chunks.push_back({z0, left});
    }
  }
  const size_t n_chunks = chunks.size();
  r->t.chunks = (uint32_t)n_chunks;
  // Two slab (and class-plane) buffers when there is more than one chunk: K1/K2 of chunk c+1 then run
  // on the producer stream while the consumer stream works through K3/K4a/K4b of chunk c.
  bool pipelined = !dense && n_chunks > 1 && !getenv("S2M_NO_CHUNK_OVERLAP");
  if (pipelined && !no_slab) {
    const size_t max_planes = (size_t)chunks[0].nzc + 1;
    if (c->slab2.ensure(max_planes * plane_bytes) != S2M_OK) { cudaGetLastError(); pipelined = false; }  // not enough memory: one buffer, no overlap
  }
  if ((st = c->ensure_slots(n_chunks + 1))) return st;
  if ((st = c->chunk_tot.ensure((n_chunks + 2) * 16))) return st;
  CUDA_TRY(cudaMemsetAsync(c->chunk_tot.p, 0, (n_chunks + 2) * 16, s));
  unsigned long long* d_tot = c->chunk_tot.as<unsigned long long>();
  // Host output regions: address ranges for the largest output this slab can have (one vertex per cell, three
  // quads per vertex; capped -- vertex ranks are 32-bit), pages are locked as the chunks report their sizes.
  {
    const unsigned long long cells = (unsigned long long)g.res[0] * g.res[1] * std::max<uint32_t>(r->nz, 1u);
    const unsigned long long vmax = std::min<unsigned long long>(cells, 0xffffffffull);
    const unsigned long long qmax = std::min<unsigned long long>(3ull * vmax, 1ull << 33);
    r->g_pos = c->lease_region(0, vmax * 12); r->g_nrm = c->lease_region(1, vmax * 12);
    r->g_key = c->lease_region(2, vmax * 8); r->g_nib = c->lease_region(3, vmax);
    r->g_quads = c->lease_region(4, qmax * r->quad_bytes());
    if (!r->g_pos || !r->g_nrm || !r->g_key || !r->g_nib || !r->g_quads)
      return fail(S2M_ERR_OOM, "mmap of the host output address ranges failed");
  }
  tr.mark("setup");

  uint64_t cand_done = 0, vert_done = 0, quad_done = 0;
  std::vector<uint32_t> dense_row;
  cudaStream_t ps = pipelined ? c->prod_stream : s;   // producer stream (K1, K2)
  if (pipelined) CUDA_TRY(cudaStreamWaitEvent(ps, c->ev[EV_BEGIN], 0));  // after the counter reset
  const unsigned cls_words = g.pitch_x / 32u;
  static const bool carry_planes = [] { const char* e = getenv("S2M_CARRY_PLANES"); return !e || atoi(e) != 0; }();
  const float* coord[3] = {nullptr, nullptr, nullptr};
  {
    unsigned bx, by;
    k1_block_shape(&bx, &by);
    if ((st = prepare_coords(c, m, g, bx, by, ps, coord))) return st;
  }
  // K1 of chunk ci into slab / class-plane buffer ci % 2 (buffer 0 when not pipelined), on the producer stream
  auto produce = [&](size_t ci) -> int {
    const Chunk ch = chunks[ci];
    const int buf = pipelined ? (int)(ci & 1) : 0;
    DevBuf& slab_buf = buf ? c->slab2 : c->slab;
    DevBuf& cls_buf = buf ? c->cls2 : c->cls;
    int st2;
    GridDev gd = g;
    float* slab = no_slab ? nullptr : slab_buf.as<float>();
    unsigned first_plane = r->z_first + ch.z0, n_planes = ch.nzc + 1;
    unsigned cw = cls_words;
    void* cls = nullptr;
    if (!from_slab) {
      if ((st2 = cls_buf.ensure((size_t)n_planes * g.rows * cls_words * 8 + 64))) return st2;
      cls = cls_buf.p;
    }
    if (pipelined && ci >= 2) CUDA_TRY(cudaStreamWaitEvent(ps, c->ev[EV_CONSUMED0 + buf], 0));  // K2 and K4a of chunk ci-2 have read this buffer
    // Consecutive chunks share one corner plane (the top of chunk ci-1 is the bottom of chunk ci): K1's plane-0 blocks
    // copy it from the previous chunk's buffers instead of evaluating it again (kernels_jit.cuh).  (Two
    // cudaMemcpyAsync in front of the launch did the same but left a 30 us gap on the producer stream, as much as
    // they saved.)  Same values either way; S2M_CARRY_PLANES=0 evaluates every plane.
    const float* carry_slab = nullptr;
    const void* carry_cls = nullptr;
    if (ci >= 1 && carry_planes) {
      const int pbuf = pipelined ? (int)((ci - 1) & 1) : 0;
      const size_t src_plane = chunks[ci - 1].nzc;
      if (slab) carry_slab = (pbuf ? c->slab2 : c->slab).as<float>() + src_plane * g.plane_stride;
      if (cls) carry_cls = static_cast<const char*>((pbuf ? c->cls2 : c->cls).p) + src_plane * (size_t)g.rows * cls_words * 8;
      if (!pipelined) { carry_slab = nullptr; carry_cls = nullptr; }   // one buffer: source and destination planes would be the same launch's
    }
    float tau_arg = tau;
    const float* coord_z = coord[2] ? coord[2] + first_plane : nullptr;
    unsigned opt = ((carry_slab || carry_cls) ? 1u : 0u) | (slab ? 2u : 0u) | (cls ? 4u : 0u);
    unsigned bx, by;
    k1_block_shape(&bx, &by);
    const unsigned gx = (g.pitch_x + 4u * bx - 1u) / (4u * bx), gy = (g.rows + by * m->k1_rows - 1u) / (by * m->k1_rows);
    // planes a thread marches through: the module's maximum where the grid is large, fewer where that would leave the
    // launch with less than ~8 waves of blocks (148 SMs x 4 resident blocks): a 256^3 grid gets 4, a 128^3 grid 1
    unsigned zpt = (unsigned)std::max<unsigned long long>(1, std::min<unsigned long long>(m->k1_zpt, (unsigned long long)gx * gy * n_planes / 4736ull));
    void* a1[] = {&gd, &slab, &first_plane, &n_planes, &tau_arg, &cls, &cw, &carry_slab, &carry_cls, &coord[0], &coord[1], &coord_z, &opt, &zpt};
    dim3 grid1(gx, gy, (n_planes + zpt - 1u) / zpt);
    {
      SPAN_BEGIN(0, ps);
      if ((st2 = launch(m->k1, grid1, dim3(bx, by, 1), ps, a1, "s2m_k1_slab"))) return st2;
      SPAN_END(ps);
    }
    CUDA_TRY(cudaEventRecord(c->ev[EV_K1DONE0 + buf], ps));
    r->t.launches += 1;
    return S2M_OK;
  };
  // K2 of chunk ci on the CONSUMER stream: it is bound by memory, K1 by instruction issue, so K2 of chunk c runs beside
  // K1 of chunk c+1 instead of between two K1 launches (1.1 ms of a 2048^3 run).  Its last block writes the chunk's
  // candidate count into mapped host memory; EV_PRODUCED is what the host waits for.
  auto classify = [&](size_t ci) -> int {
    const Chunk ch = chunks[ci];
    const int buf = pipelined ? (int)(ci & 1) : 0;
    float* slab = no_slab ? nullptr : (buf ? c->slab2 : c->slab).as<float>();
    void* cls = from_slab ? nullptr : (buf ? c->cls2 : c->cls).p;
    if (pipelined) CUDA_TRY(cudaStreamWaitEvent(s, c->ev[EV_K1DONE0 + buf], 0));
    unsigned long long* cnt = d_cnt + (buf ? C_CHUNK_CAND1 : C_CHUNK_CAND0);
    CUDA_TRY(cudaMemsetAsync(cnt, 0, 16, s));   // the count and K2's block-completion counter
    S2mK2Args a2{};
    a2.slab = slab; a2.pitch_x = g.pitch_x; a2.plane_stride = g.plane_stride;
    a2.res_x = g.res[0]; a2.res_y = g.res[1]; a2.nz_chunk = ch.nzc; a2.tau = tau;
    a2.cand_mask = c->cand_mask.as<uint32_t>() + words_per_slice * ch.z0; a2.words_x = r->words_x; a2.total = cnt;
    a2.cls = cls; a2.cls_words = cls_words;
    a2.done = reinterpret_cast<unsigned*>(cnt + 1); a2.host_total = c->h_counters + buf;
    a2.seg_count = c->seg_count.as<uint32_t>() + segs_per_slice * ch.z0;
    {
      SPAN_BEGIN(1, s);
      int e2 = from_slab ? s2m_launch_k2(&a2, s) : s2m_launch_k2_bits(&a2, s);
      if (!e2 && from_slab) e2 = s2m_launch_seg_count(a2.cand_mask, (unsigned long long)g.res[1] * ch.nzc, r->words_x, a2.seg_count, s);
      if (e2) return fail(S2M_ERR_CUDA, std::string("k2_classify launch: ") + cudaGetErrorString((cudaError_t)e2));
      SPAN_END(s);
    }
    CUDA_TRY(cudaEventRecord(c->ev[EV_PRODUCED0 + buf], s));
    r->t.launches += 1;
    return S2M_OK;
  };
  if (!dense && n_chunks && (st = produce(0))) return st;
  for (size_t ci = 0; ci < n_chunks; ++ci) {
    const Chunk ch = chunks[ci];
    const int buf = pipelined ? (int)(ci & 1) : 0;
    const unsigned long long chunk_words = words_per_slice * ch.nzc;
    uint32_t* mask_chunk = c->cand_mask.as<uint32_t>() + words_per_slice * ch.z0;
    unsigned slab_first_plane = 0, slab_n_planes = 0;
    uint64_t n_cand = 0;
    if (!dense) {
      // the next chunk's K1 is queued before this chunk's count is waited for
      if (pipelined && ci + 1 < n_chunks && (st = produce(ci + 1))) return st;
      if ((st = classify(ci))) return st;
      if (!no_slab) { slab_first_plane = r->z_first + ch.z0; slab_n_planes = ch.nzc + 1; }
      // ---- [count] candidates of this chunk: K2's last block wrote it into mapped host memory
      CUDA_TRY(cudaEventSynchronize(c->ev[EV_PRODUCED0 + buf]));
      n_cand = c->h_counters[buf];
    } else {
      // reference-cost mode: every cell of the chunk is a candidate
      if (dense_row.empty()) {
        dense_row.assign(r->words_x, 0xffffffffu);
        if (g.res[0] % 32u) dense_row.back() = (1u << (g.res[0] % 32u)) - 1u;
      }
      std::vector<uint32_t> host(chunk_words);
      for (unsigned long long i = 0; i < chunk_words; i += r->words_x) memcpy(&host[i], dense_row.data(), r->words_x * 4);
      CUDA_TRY(cudaMemcpyAsync(mask_chunk, host.data(), chunk_words * 4, cudaMemcpyHostToDevice, s));
      if (int e = s2m_launch_seg_count(mask_chunk, (unsigned long long)g.res[1] * ch.nzc, r->words_x, c->seg_count.as<uint32_t>() + segs_per_slice * ch.z0, s))
        return fail(S2M_ERR_CUDA, std::string("k_seg_count launch: ") + cudaGetErrorString((cudaError_t)e));
      CUDA_TRY(cudaStreamSynchronize(s));
      n_cand = (unsigned long long)g.res[0] * g.res[1] * ch.nzc;
    }
    const uint64_t cand_total = cand_done + n_cand;
    if (cand_total >= 0xffffffffull) return fail(S2M_ERR_UNSUPPORTED, "more than 2^32-1 candidate cells in one slab; split it with z_begin/z_end");
    const unsigned k3_tiles = s2m_k3_tiles(chunk_words, r->words_x);
    const unsigned k4_tiles = (unsigned)std::max<uint64_t>(1, (n_cand + 127) / 128);   // >= 1: an empty chunk still carries the totals forward
    const unsigned k4b_tiles = s2m_k4b_tiles(n_cand);
    const size_t status_words = (size_t)k3_tiles + k4_tiles + k4b_tiles + 16;
    const size_t qb = r->quad_bytes();
    // vertices <= candidates, quads <= 3 per vertex: exact upper bounds from counts the host already has
    if ((st = c->status.ensure(status_words * 8))) return st;
    if ((st = c->cand_key.ensure_preserve((cand_total + 1) * 8, cand_done * 8, s))) return st;
    if ((st = c->cand_vrank.ensure_preserve((cand_total + 1) * 4, cand_done * 4, s))) return st;
    if ((st = c->v_pos.ensure_preserve((vert_done + n_cand + 1) * 12, vert_done * 12, s))) return st;
    if ((st = c->v_nrm.ensure_preserve((vert_done + n_cand + 1) * 12, vert_done * 12, s))) return st;
    if ((st = c->v_key.ensure_preserve((vert_done + n_cand + 1) * 8, vert_done * 8, s))) return st;
    if ((st = c->v_nib.ensure_preserve(vert_done + n_cand + 16, vert_done, s))) return st;
    if ((st = c->quads.ensure_preserve((size_t)(quad_done + 3 * n_cand) * qb + 64, (size_t)quad_done * qb, s))) return st;
    CUDA_TRY(cudaMemsetAsync(c->status.p, 0, status_words * 8, s));
    unsigned long long* status = c->status.as<unsigned long long>();
    unsigned* tickets = reinterpret_cast<unsigned*>(status);  // words 0..1: K3 ticket, K4a ticket, K4b ticket, K4b completion counter; 2..: tile status
    unsigned long long* tot_prev = d_tot + 2 * ci;
    unsigned long long* tot_cur = d_tot + 2 * (ci + 1);
    // ---- K3
    {
      S2mK3Args a3{};
      a3.cand_mask = mask_chunk; a3.seg_count = c->seg_count.as<uint32_t>() + segs_per_slice * ch.z0; a3.n_words = chunk_words; a3.words_x = r->words_x; a3.res_y = g.res[1];
      a3.z_offset = r->z_first + ch.z0; a3.word_prefix = c->word_prefix.as<uint32_t>() + words_per_slice * ch.z0;
      a3.cand_key = c->cand_key.as<unsigned long long>(); a3.base = cand_done;
      a3.status = status + 2; a3.ticket = tickets;
      SPAN_BEGIN(2, s);
      int e3 = s2m_launch_k3(&a3, s);
      if (e3) return fail(S2M_ERR_CUDA, std::string("k3_compact launch: ") + cudaGetErrorString((cudaError_t)e3));
      SPAN_END(s);
    }
    // ---- K4a
    {
      GridDev gd = g;
      const unsigned long long* ck = c->cand_key.as<unsigned long long>() + cand_done;
      unsigned long long nc = n_cand;
      const unsigned long long* vbase = tot_prev;   // vertices emitted by earlier chunks, read on the device
      unsigned label_add = r->label_add, halo_below = r->halo ? (r->z_first + 1u) : 0u;
      unsigned want_normals = ((p->flags & S2M_MESH_NO_NORMALS) ? 0u : 1u) | ((p->flags & S2M_MESH_CONSISTENT_CORNERS) ? 2u : 0u);  // K4a's mode bits
      SlabViewDev sv{no_slab || dense ? nullptr : (buf ? c->slab2 : c->slab).as<float>(), slab_first_plane, slab_n_planes};
      VertexOutDev vo{c->v_pos.as<float>(), c->v_nrm.as<float>(), c->v_key.as<unsigned long long>(), c->v_nib.as<unsigned char>(),
                      c->cand_vrank.as<unsigned>() + cand_done, status + 2 + k3_tiles, tickets + 1, tot_cur, d_cnt + C_NHALO};
      void* a4[] = {&gd, &ck, &nc, &vbase, &label_add, &halo_below, &want_normals, &sv, &vo};
      SPAN_BEGIN(3, s);
      // (capping K4a's blocks per SM with unused dynamic shared memory, to keep K1 blocks of the next chunk
      // resident beside them, was measured: K4a 4.3 -> 6.4 ms and the run got 2 ms slower)
      if ((st = launch(m->k4, dim3(k4_tiles), dim3(128), s, a4, "s2m_k4_vertices"))) return st;
      SPAN_END(s);
    }
    if (pipelined) CUDA_TRY(cudaEventRecord(c->ev[EV_CONSUMED0 + buf], s));  // slab buffer `buf` may be overwritten
    // ---- K4b: this chunk's quads, slab-relative indices (local - n_halo); its last block publishes the totals
    {
      S2mK4bArgs a{};
      a.vert_key = c->v_key.as<unsigned long long>(); a.vert_nibble = c->v_nib.as<unsigned char>();
      a.max_vertices = n_cand; a.tot_prev = tot_prev; a.tot_cur = tot_cur; a.n_halo = d_cnt + C_NHALO;
      a.cand_mask = c->cand_mask.as<uint32_t>(); a.word_prefix = c->word_prefix.as<uint32_t>(); a.cand_vrank = c->cand_vrank.as<uint32_t>();
      a.words_x = r->words_x; a.res_y = g.res[1]; a.z_first = r->z_first; a.label_add = r->label_add;
      a.index_add = 0;
      a.quads = c->quads.as<unsigned long long>(); a.quads32 = r->quads_u32() ? c->quads.as<unsigned>() : nullptr;
      a.status = status + 2 + k3_tiles + k4_tiles; a.ticket = tickets + 2; a.n_invalid = d_cnt + C_NINVALID;
      a.done = tickets + 3; a.host_slot = c->h_slots + 8 * ci;
      if (p->flags & S2M_MESH_KEEP_INVALID) {
        if ((st = c->invalid.ensure(kInvalidCapacity * 48))) return st;
        a.invalid_records = c->invalid.as<unsigned long long>(); a.invalid_cursor = d_cnt + C_INVALID_CURSOR; a.invalid_capacity = kInvalidCapacity;
      }
      SPAN_BEGIN(4, s);
      int e = s2m_launch_k4b(&a, s);
      if (e) return fail(S2M_ERR_CUDA, std::string("k4_quads launch: ") + cudaGetErrorString((cudaError_t)e));
      SPAN_END(s);
    }
    r->t.launches += 3;
    const size_t ev_done = take_event(c, r);
    CUDA_TRY(cudaEventRecord(c->ev_pool[ev_done], s));
    if (ci + 1 == n_chunks) CUDA_TRY(cudaEventRecord(c->ev[1], s));  // device_ms: every kernel has finished
    if (!pipelined && !dense && ci + 1 < n_chunks && (st = produce(ci + 1))) return st;
    // ---- [count] what this chunk added; then its copies, while the next chunk computes
    CUDA_TRY(cudaEventSynchronize(c->ev_pool[ev_done]));
    const unsigned long long* slot = c->h_slots + 8 * ci;
    const uint64_t vert_total = slot[0], quad_total = slot[1];
    r->n_halo = slot[2]; r->n_invalid = slot[3]; r->n_invalid_records = std::min<uint64_t>(slot[4], kInvalidCapacity);
    CUDA_TRY(cudaStreamWaitEvent(c->copy_stream, c->ev_pool[ev_done], 0));
    if (vert_total > std::max<uint64_t>(r->copied_v, r->n_halo) || quad_total > r->copied_q) {
      SPAN_BEGIN(5, c->copy_stream);
      if ((st = copy_out(c, r, c->copy_stream, vert_total, quad_total))) return st;
      SPAN_END(c->copy_stream);
    }
    cand_done = cand_total; vert_done = vert_total; quad_done = quad_total;
    tr.mark("chunk done");
  }
  if (n_chunks == 0) CUDA_TRY(cudaEventRecord(c->ev[1], s));
  r->n_cand = cand_done;
  r->n_vert_total = vert_done;
  r->n_quads = quad_done;
  const uint64_t n_own = r->n_vert_total - r->n_halo;
  if (p->flags & S2M_MESH_NO_NORMALS) memset(r->g_nrm->base, 0, n_own * 12);   // the result hands out zeros (plain pages: nothing is copied into them)
  if (p->flags & S2M_MESH_KEEP_CANDIDATES) {
    r->h_cand = (uint64_t*)c->lease_pinned(r->n_cand * 8);
    if (!r->h_cand) return fail(S2M_ERR_OOM, "cudaHostAlloc for candidate list failed");
    if (r->n_cand) CUDA_TRY(cudaMemcpyAsync(r->h_cand, c->cand_key.p, r->n_cand * 8, cudaMemcpyDeviceToHost, c->copy_stream));
  }
  if (r->n_halo) {  // the halo slice's positions, so that this slab's triangles can be written without the slab below
    r->h_halo_pos = (float*)c->lease_pinned(r->n_halo * 12);
    if (!r->h_halo_pos) return fail(S2M_ERR_OOM, "cudaHostAlloc for halo positions failed");
    CUDA_TRY(cudaMemcpyAsync(r->h_halo_pos, c->v_pos.p, r->n_halo * 12, cudaMemcpyDeviceToHost, c->copy_stream));
  }
  if ((p->flags & S2M_MESH_KEEP_INVALID) && r->n_invalid_records) {
    r->h_invalid = (uint64_t*)c->lease_pinned(r->n_invalid_records * 48 + 48);
    if (!r->h_invalid) return fail(S2M_ERR_OOM, "cudaHostAlloc for the invalid-quad list failed");
    CUDA_TRY(cudaMemcpyAsync(r->h_invalid, c->invalid.p, r->n_invalid_records * 48, cudaMemcpyDeviceToHost, c->copy_stream));
  }
  CUDA_TRY(cudaEventRecord(c->ev[2], c->copy_stream));   // total_ms: everything resident in pinned host memory
  tr.mark("begin done");
  *out = rp.release();
  return S2M_OK;
}

// index += base over n indices, on a few host threads (the non-relative contract of s2m_mesh_finish)
template <class T>
void add_base(T* q, uint64_t n, T base) {
  const unsigned hw = std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
  const unsigned nt = n < (1u << 18) ? 1u : hw;
  auto work = [=](unsigned k) {
    const uint64_t a = n * k / nt, b = n * (k + 1) / nt;
    for (uint64_t i = a; i < b; ++i) q[i] = (T)(q[i] + base);
  };
  std::vector<std::thread> th;
  for (unsigned k = 1; k < nt; ++k) {
    try { th.emplace_back(work, k); } catch (const std::system_error&) { work(k); }
  }
  work(0);
  for (auto& t : th) t.join();
}

}  // namespace

extern "C" int s2m_mesh_begin(s2m_ctx* c, s2m_module* m, const s2m_mesh_params* p, s2m_result** out) {
  return mesh_begin_impl(c, m, p, out);
}

extern "C" int s2m_mesh_finish(s2m_result* r, int64_t global_vertex_base) {
  if (!r || !r->ctx) return fail(S2M_ERR_INVALID_ARG, "s2m_mesh_finish: NULL result");
  if (r->finished) return fail(S2M_ERR_STATE, "s2m_mesh_finish called twice");
  s2m_ctx* c = r->ctx;
  Trace tr;
  CUDA_TRY(cudaSetDevice(c->device));
  r->global_base = global_vertex_base;
  if (r->quads_u32() && (global_vertex_base < 0 || (uint64_t)global_vertex_base + (r->n_vert_total - r->n_halo) > 0xffffffffull))
    return fail(S2M_ERR_UNSUPPORTED, "S2M_MESH_QUADS_U32: vertex indices of this slab do not fit 32 bits");
  // everything was computed and queued for copying by begin(); what is left is to wait for the copies
  CUDA_TRY(cudaStreamSynchronize(c->copy_stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  tr.mark("finish: copies done");
  if (r->h_invalid && r->n_invalid_records) {
    // the device appends in completion order; the reference reports them in (vertex, edge) order
    struct Rec { uint64_t v[6]; };
    Rec* recs = reinterpret_cast<Rec*>(r->h_invalid);
    std::sort(recs, recs + r->n_invalid_records, [](const Rec& x, const Rec& y) { return x.v[0] != y.v[0] ? x.v[0] < y.v[0] : x.v[1] < y.v[1]; });
  }
  if (!r->relative() && global_vertex_base != 0) {
    // the default contract: quads (and invalid records) carry GLOBAL indices.  The device wrote slab-relative ones
    // (local - n_halo) so that nothing had to wait for the base; S2M_MESH_RELATIVE_QUADS keeps them that way.
    if (r->quads_u32()) add_base<uint32_t>(r->host<uint32_t>(r->g_quads), 4 * r->n_quads, (uint32_t)global_vertex_base);
    else add_base<uint64_t>(r->host<uint64_t>(r->g_quads), 4 * r->n_quads, (uint64_t)global_vertex_base);
    for (uint64_t i = 0; i < r->n_invalid_records && r->h_invalid; ++i)
      for (int t = 2; t < 6; ++t)
        if (r->h_invalid[6 * i + t] != ~0ull) r->h_invalid[6 * i + t] += (uint64_t)global_vertex_base;
    tr.mark("finish: base added");
  }
  r->t.host_wall_ms = now_ms() - r->wall0;
  finalize_timings(c, r);
  tr.mark("finish: timings read");
  r->finished = true;
  c->busy = false;
  return S2M_OK;
}

extern "C" int s2m_mesh_run(s2m_ctx* c, s2m_module* m, const s2m_mesh_params* p, s2m_result** out) {
  int st = mesh_begin_impl(c, m, p, out);
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
  o->positions = r->host<float>(r->g_pos); o->normals = r->host<float>(r->g_nrm);
  o->cell_keys = r->host<uint64_t>(r->g_key); o->sign_nibbles = r->host<uint8_t>(r->g_nib);
  o->quads = r->quads_u32() ? nullptr : r->host<uint64_t>(r->g_quads); o->quads32 = r->quads_u32() ? r->host<uint32_t>(r->g_quads) : nullptr;
  o->quad_index_add = r->relative() ? r->global_base : 0;
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
extern "C" uint64_t s2m_module_uid(const s2m_module* m) { return m ? m->uid : 0; }
extern "C" int s2m_module_prefers_no_slab(const s2m_module* m) { return m && m->slab_free_default ? 1 : 0; }

extern "C" int s2m_measure_fp32_peak(s2m_ctx* c, double out_tflops[3]) {
  if (!c || !out_tflops) return fail(S2M_ERR_INVALID_ARG, "NULL argument");
  CUDA_TRY(cudaSetDevice(c->device));
  DevBuf sink;
  int st = sink.ensure(64);
  if (st) return st;
  cudaStream_t s = c->stream;
  const int blocks = c->prop.multiProcessorCount * 8, iters = 1 << 15;   // 8 x 256 threads per SM, 2^18 FMAs per thread: ~2 ms
  const double flops = 2.0 * 8.0 * iters * 256.0 * blocks;
  for (int mode = 0; mode < 3; ++mode) {
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {   // rep 0 warms up
      CUDA_TRY(cudaEventRecord(c->ev[8], s));
      int e = s2m_launch_fp32_probe(mode, blocks, iters, sink.as<float>(), s);
      if (e) { sink.release(); return fail(S2M_ERR_CUDA, std::string("fp32 probe launch: ") + cudaGetErrorString((cudaError_t)e)); }
      CUDA_TRY(cudaEventRecord(c->ev[9], s));
      CUDA_TRY(cudaEventSynchronize(c->ev[9]));
      float ms = 0;
      cudaEventElapsedTime(&ms, c->ev[8], c->ev[9]);
      if (rep && ms > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    out_tflops[mode] = best;
  }
  sink.release();
  return S2M_OK;
}

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
  const float* carry_slab = nullptr;
  const void* carry_cls = nullptr;
  unsigned bx, by;
  k1_block_shape(&bx, &by);
  const float* coord[3];
  if ((st = prepare_coords(c, m, g, bx, by, c->stream, coord))) return st;
  const float* coord_z = coord[2] ? coord[2] + first_plane : nullptr;
  unsigned opt = 2u;   // the slab only
  unsigned zpt = 1;
  void* a1[] = {&g, &slab, &first_plane, &n_planes, &tau_arg, &cls, &cls_words, &carry_slab, &carry_cls, &coord[0], &coord[1], &coord_z, &opt, &zpt};
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

// Relative cost of the z-bands of a grid, for cutting it into slabs of equal work: K1 itself -- the module's own kernel,
// launch shape and output buffers -- on ONE corner plane in the middle of every band, timed with events.  (Until round 2
// a lattice of scalar evaluations timed with clock64, s2m_k_cost_probe: once K1's per-thread overhead had been cut it
// overrated the cheap bands, and the first partition needed two refinements to recover.  S2M_COST_PROBE=lattice keeps it.)
static int cost_probe_lattice(s2m_ctx* c, s2m_module* m, const GridDev& g, uint32_t planes, double* cost_out) {
  DevBuf buf;
  int st;
  if ((st = buf.ensure((size_t)planes * 8 + 64))) return st;
  cudaStream_t s = c->stream;
  cudaError_t e = cudaMemsetAsync(buf.p, 0, (size_t)planes * 8 + 64, s);
  unsigned probe = 128, pl = planes;
  unsigned long long* cyc = buf.as<unsigned long long>();
  float* sink = reinterpret_cast<float*>(cyc + planes);
  GridDev gd = g;
  void* args[] = {&gd, &probe, &pl, &cyc, &sink};
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

extern "C" int s2m_cost_probe(s2m_ctx* c, s2m_module* m, const s2m_mesh_params* p, uint32_t planes, double* cost_out) {
  if (!c || !m || !p || !cost_out || planes == 0) return fail(S2M_ERR_INVALID_ARG, "bad argument");
  if (!m->k1 || m->ctx != c) return fail(S2M_ERR_STATE, "module was not compiled for this ctx");
  if (c->busy) return fail(S2M_ERR_STATE, "ctx is busy");
  GridDev g;
  int st = make_grid(p, &g);
  if (st) return st;
  CUDA_TRY(cudaSetDevice(c->device));
  if (const char* e = getenv("S2M_COST_PROBE")) if (!strcmp(e, "lattice")) return cost_probe_lattice(c, m, g, planes, cost_out);
  cudaStream_t s = c->stream;
  const bool no_slab = ((p->flags & S2M_MESH_NO_SLAB) || m->slab_free_default) && !(p->flags & (S2M_MESH_CLASSIFY_FROM_SLAB | S2M_MESH_EXACT_DENSE));
  const unsigned cls_words = g.pitch_x / 32u;
  DevBuf slab_buf, cls_buf;
  if (!no_slab && (st = slab_buf.ensure(g.plane_stride * 4))) return st;
  if ((st = cls_buf.ensure((size_t)g.rows * cls_words * 8 + 64))) return st;
  unsigned bx, by;
  k1_block_shape(&bx, &by);
  const float* coord[3];
  if ((st = prepare_coords(c, m, g, bx, by, s, coord))) return st;
  std::vector<cudaEvent_t> ev(planes + 1, nullptr);
  cudaError_t e = cudaSuccess;
  for (auto& x : ev) if (e == cudaSuccess) e = cudaEventCreate(&x);
  const dim3 grid((g.pitch_x + 4u * bx - 1u) / (4u * bx), (g.rows + by * m->k1_rows - 1u) / (by * m->k1_rows), 1), block(bx, by, 1);
  float* slab = no_slab ? nullptr : slab_buf.as<float>();
  void* cls = cls_buf.p;
  unsigned n_planes = 1, cw = cls_words, opt = (slab ? 2u : 0u) | 4u, zpt = 1;
  float tau_arg = 0.0f;
  const float* carry_slab = nullptr;
  const void* carry_cls = nullptr;
  GridDev gd = g;
  for (uint32_t i = 0; i <= planes && e == cudaSuccess && !st; ++i) {   // launch 0 warms the instruction cache up and is not timed
    const uint32_t band = i == 0 ? 0u : i - 1u;
    unsigned first_plane = (unsigned)std::min<unsigned long long>(g.res[2], ((2ull * band + 1ull) * (g.res[2] + 1ull)) / (2ull * planes));
    const float* coord_z = coord[2] ? coord[2] + first_plane : nullptr;
    void* a1[] = {&gd, &slab, &first_plane, &n_planes, &tau_arg, &cls, &cw, &carry_slab, &carry_cls, &coord[0], &coord[1], &coord_z, &opt, &zpt};
    st = launch(m->k1, grid, block, s, a1, "s2m_k1_slab");
    if (!st) e = cudaEventRecord(ev[i], s);
  }
  if (e == cudaSuccess && !st) e = cudaStreamSynchronize(s);
  for (uint32_t i = 0; i < planes && e == cudaSuccess && !st; ++i) {
    float ms = 0.0f;
    e = cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
    cost_out[i] = (double)ms;
  }
  for (auto& x : ev) if (x) cudaEventDestroy(x);
  slab_buf.release(); cls_buf.release();
  if (st) return st;
  if (e != cudaSuccess) return fail(S2M_ERR_CUDA, std::string("s2m_cost_probe: ") + cudaGetErrorString(e));
  return S2M_OK;
}
