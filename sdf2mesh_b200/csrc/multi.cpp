// multi.cpp -- one process, N GPUs: z-slab meshing behind the C ABI (include/sdf2mesh_b200.h, s2m_multi_*).
//
// The reference has one adapter, one device, one queue (/root/reference/src/bin/sdf2mesh/main.rs:180-196) and one
// slice loop (:298-356).  Here the grid is cut into contiguous z-slabs, one per GPU: every GPU
// recomputes the one slice below its slab as halo, meshes its slab with s2m_mesh_begin (K1 ... K4b, quads with
// slab-relative indices), and the only exchange is ONE all-gather of the per-slab vertex counts -- ncclAllGather on a
// communicator from ncclCommInitAll, 8 bytes per rank over NVLink -- whose exclusive prefix is each slab's global
// vertex base (s2m_mesh_finish).  Each GPU is driven by its own host thread (kept alive between runs) with its own
// s2m_ctx; slab boundaries are balanced from a coarse cost probe (s2m_cost_probe) and refined once from measured times.
//
// libnccl.so.2 is dlopen'ed (the library must load on hosts without it); S2M_MULTI_NO_NCCL exchanges the counts through
// host memory instead -- kept as a switch so that both latencies can be reported.
//
// Built only on the public C ABI plus the CUDA runtime: nothing here reaches into the engine's internals.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "common.h"

using s2m_internal::fail;

// ------------------------------------------------------------------ slab boundaries (host arithmetic, no device)
// Boundaries b[0..world] over n_slices z-slices, strictly increasing, slab g getting ~1/world of the cost described by
// `cost` (relative cost of n_cost equal-thickness z bands; NULL / all-zero = equal thickness).  SURVEY.md H4: the
// mandelbulb's work sits in |p| <= 2, equal-thickness slabs starve the outer GPUs.
extern "C" int s2m_partition_slices(uint32_t n_slices, int world, const double* cost, int n_cost, uint32_t* bounds_out) {
  if (world < 1 || !bounds_out) return fail(S2M_ERR_INVALID_ARG, "s2m_partition_slices: bad argument");
  if (n_slices < (uint32_t)world)
    return fail(S2M_ERR_INVALID_ARG, std::to_string(n_slices) + " z-slices cannot be split over " + std::to_string(world) + " slabs (every slab needs at least one)");
  double sum = 0.0;
  bool usable = cost && n_cost > 0;
  for (int i = 0; usable && i < n_cost; ++i) { if (!std::isfinite(cost[i]) || cost[i] < 0) usable = false; else sum += cost[i]; }
  std::vector<double> b((size_t)world + 1);
  if (!usable || !(sum > 0.0)) {
    for (int g = 0; g <= world; ++g) b[(size_t)g] = std::nearbyint((double)n_slices * g / world);
  } else {
    double cmax = 0.0;
    for (int i = 0; i < n_cost; ++i) cmax = std::max(cmax, cost[i]);
    std::vector<double> cum((size_t)n_cost + 1, 0.0);
    for (int i = 0; i < n_cost; ++i) cum[(size_t)i + 1] = cum[(size_t)i] + std::max(cost[i], cmax * 1e-3);  // every band costs something
    for (int g = 0; g <= world; ++g) {
      const double target = cum.back() * g / world;
      int k = (int)(std::upper_bound(cum.begin(), cum.end(), target) - cum.begin()) - 1;
      k = std::max(0, std::min(k, n_cost - 1));
      const double frac = cum[(size_t)k + 1] > cum[(size_t)k] ? (target - cum[(size_t)k]) / (cum[(size_t)k + 1] - cum[(size_t)k]) : 0.0;
      b[(size_t)g] = std::nearbyint(((double)k + std::min(1.0, std::max(0.0, frac))) * n_slices / n_cost);
    }
  }
  bounds_out[0] = 0;
  for (int g = 1; g <= world; ++g) {
    const long long lo = (long long)bounds_out[g - 1] + 1, hi = (long long)n_slices - (world - g);
    bounds_out[g] = (uint32_t)std::min(hi, std::max(lo, (long long)b[(size_t)g]));
  }
  bounds_out[world] = n_slices;
  return S2M_OK;
}

// Refine boundaries from the time every slab actually took: the cost of slice z is shape(z) * corr_g inside slab g, with
// shape the coarse profile (uniform if NULL) and corr_g = seconds[g] / sum(shape over the slab); the new boundaries cut
// the cumulative model cost into equal parts.  Returns the old boundaries if a time is missing or not positive.
extern "C" int s2m_rebalance_slices(const uint32_t* bounds, int world, const double* seconds, const double* cost, int n_cost, uint32_t* bounds_out) {
  if (!bounds || !seconds || !bounds_out || world < 1) return fail(S2M_ERR_INVALID_ARG, "s2m_rebalance_slices: bad argument");
  const uint32_t n = bounds[world];
  bool ok = world >= 2;
  for (int g = 0; g < world && ok; ++g) ok = bounds[g + 1] > bounds[g] && std::isfinite(seconds[g]) && seconds[g] > 0;
  if (!ok) { for (int g = 0; g <= world; ++g) bounds_out[g] = bounds[g]; return S2M_OK; }
  std::vector<double> dens(n, 1.0);
  double csum = 0.0, cmax = 0.0;
  for (int i = 0; cost && i < n_cost; ++i) { csum += cost[i]; cmax = std::max(cmax, cost[i]); }
  if (cost && n_cost > 0 && std::isfinite(csum) && csum > 0) {
    for (uint32_t z = 0; z < n; ++z) {   // piecewise-linear interpolation between band centres
      const double pos = ((double)z + 0.5) * n_cost / n - 0.5;
      const int k0 = (int)std::floor(pos);
      const double t = pos - k0;
      const double a = std::max(cost[std::max(0, std::min(n_cost - 1, k0))], cmax * 1e-3), bb = std::max(cost[std::max(0, std::min(n_cost - 1, k0 + 1))], cmax * 1e-3);
      dens[z] = a + (bb - a) * t;
    }
  }
  for (int g = 0; g < world; ++g) {
    double sl = 0.0;
    for (uint32_t z = bounds[g]; z < bounds[g + 1]; ++z) sl += dens[z];
    for (uint32_t z = bounds[g]; z < bounds[g + 1]; ++z) dens[z] *= seconds[g] / sl;
  }
  std::vector<double> cum((size_t)n + 1, 0.0);
  for (uint32_t z = 0; z < n; ++z) cum[(size_t)z + 1] = cum[z] + dens[z];
  bounds_out[0] = 0;
  for (int g = 1; g <= world; ++g) {
    const double target = cum.back() * g / world;
    size_t k = (size_t)(std::upper_bound(cum.begin(), cum.end(), target) - cum.begin());
    k = std::max<size_t>(1, std::min<size_t>(k, n)) - 1;
    const double frac = cum[k + 1] > cum[k] ? (target - cum[k]) / (cum[k + 1] - cum[k]) : 0.0;
    const long long want = (long long)std::nearbyint((double)k + frac);
    const long long lo = (long long)bounds_out[g - 1] + 1, hi = (long long)n - (world - g);
    bounds_out[g] = (uint32_t)std::min(hi, std::max(lo, want));
  }
  bounds_out[world] = n;
  return S2M_OK;
}

// ------------------------------------------------------------------ NCCL (dlopen'ed)
namespace {
typedef struct ncclComm* ncclComm_t_;
struct NcclApi {
  void* handle = nullptr;
  int (*CommInitAll)(ncclComm_t_*, int, const int*) = nullptr;
  int (*CommDestroy)(ncclComm_t_) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t_, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
  bool ok = false;
  std::string why;
};
constexpr int kNcclUint64 = 5;  // ncclDataType_t::ncclUint64 (nccl.h; stable across NCCL 2.x)
NcclApi& nccl() {
  static NcclApi a;
  static std::once_flag once;
  std::call_once(once, [] {
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      a.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (a.handle) break;
    }
    if (!a.handle) { a.why = std::string("cannot load libnccl.so.2: ") + dlerror(); return; }
#define S2M_NSYM(field, sym) \
  a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.handle, sym)); \
  if (!a.field) { a.why = std::string("libnccl lacks ") + sym; return; }
    S2M_NSYM(CommInitAll, "ncclCommInitAll")
    S2M_NSYM(CommDestroy, "ncclCommDestroy")
    S2M_NSYM(AllGather, "ncclAllGather")
    S2M_NSYM(GetErrorString, "ncclGetErrorString")
    S2M_NSYM(GetVersion, "ncclGetVersion")
#undef S2M_NSYM
    a.ok = true;
  });
  return a;
}

double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// a reusable barrier for the host-memory exchange (no NCCL)
struct SpinBarrier {
  std::atomic<unsigned> count{0}, generation{0};
  void wait(unsigned n) {
    const unsigned gen = generation.load(std::memory_order_acquire);
    if (count.fetch_add(1, std::memory_order_acq_rel) + 1 == n) { count.store(0, std::memory_order_relaxed); generation.fetch_add(1, std::memory_order_release); }
    else while (generation.load(std::memory_order_acquire) == gen) std::this_thread::yield();
  }
};
}  // namespace

struct s2m_multi {
  int n = 0;
  uint32_t flags = 0;
  std::vector<int> ordinals;
  std::vector<s2m_ctx*> ctx;
  std::vector<ncclComm_t_> comm;
  std::vector<unsigned long long*> d_mine, d_all;      // per device: this slab's count, everybody's counts
  std::vector<unsigned long long*> h_mine;             // per device: pinned staging word
  std::vector<cudaStream_t> xs;                        // per device: the stream the exchange runs on
  bool use_nccl = false;
  int nccl_version = 0;
  // modules instantiated per device, keyed by s2m_module_uid of the compiled module they came from (not by its address:
  // the caller may free it and compile another one that lands at the same address); the oldest of more than 8 is dropped
  std::map<uint64_t, std::vector<s2m_module*>> modules;
  std::vector<uint64_t> module_order;
  // partition state for the last (module, grid) seen
  uint64_t part_module = 0;
  s2m_mesh_params part_params{};
  std::vector<uint32_t> bounds;
  std::vector<double> cost;
  int runs_on_partition = 0;
  // per-run scratch
  std::vector<uint64_t> counts;
  SpinBarrier barrier;
  s2m_multi_timings last{};
  // worker threads
  struct Worker {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::function<void()> job;
    bool has_job = false, done = false, quit = false;
  };
  std::vector<std::unique_ptr<Worker>> workers;

  void run_on_all(const std::function<void(int)>& f) {
    for (int k = 0; k < n; ++k) {
      Worker& w = *workers[(size_t)k];
      std::lock_guard<std::mutex> lk(w.mu);
      w.job = [&f, k] { f(k); };
      w.has_job = true; w.done = false;
      w.cv.notify_all();
    }
    for (int k = 0; k < n; ++k) {
      Worker& w = *workers[(size_t)k];
      std::unique_lock<std::mutex> lk(w.mu);
      w.cv.wait(lk, [&w] { return w.done; });
    }
  }
};

namespace {
void worker_loop(s2m_multi::Worker* w) {
  for (;;) {
    std::function<void()> job;
    {
      std::unique_lock<std::mutex> lk(w->mu);
      w->cv.wait(lk, [w] { return w->has_job || w->quit; });
      if (w->quit) return;
      job = std::move(w->job);
      w->has_job = false;
    }
    job();
    {
      std::lock_guard<std::mutex> lk(w->mu);
      w->done = true;
    }
    w->cv.notify_all();
  }
}
bool same_grid(const s2m_mesh_params& a, const s2m_mesh_params& b) {
  return memcmp(a.bb_min, b.bb_min, sizeof a.bb_min) == 0 && memcmp(a.bb_max, b.bb_max, sizeof a.bb_max) == 0 &&
         memcmp(a.dims, b.dims, sizeof a.dims) == 0 && ((a.flags ^ b.flags) & S2M_MESH_ALL_SLICES) == 0;
}
}  // namespace

extern "C" void s2m_multi_destroy(s2m_multi* mc) {
  if (!mc) return;
  for (auto& w : mc->workers) {
    { std::lock_guard<std::mutex> lk(w->mu); w->quit = true; }
    w->cv.notify_all();
    if (w->th.joinable()) w->th.join();
  }
  for (auto& kv : mc->modules) for (s2m_module* m : kv.second) if (m) s2m_module_free(m);
  for (int k = 0; k < (int)mc->ctx.size(); ++k) {
    if (!mc->ctx[(size_t)k]) continue;
    cudaSetDevice(mc->ordinals[(size_t)k]);
    if (k < (int)mc->comm.size() && mc->comm[(size_t)k]) nccl().CommDestroy(mc->comm[(size_t)k]);
    if (k < (int)mc->d_mine.size() && mc->d_mine[(size_t)k]) cudaFree(mc->d_mine[(size_t)k]);
    if (k < (int)mc->d_all.size() && mc->d_all[(size_t)k]) cudaFree(mc->d_all[(size_t)k]);
    if (k < (int)mc->h_mine.size() && mc->h_mine[(size_t)k]) cudaFreeHost(mc->h_mine[(size_t)k]);
    if (k < (int)mc->xs.size() && mc->xs[(size_t)k]) cudaStreamDestroy(mc->xs[(size_t)k]);
    s2m_ctx_destroy(mc->ctx[(size_t)k]);
  }
  delete mc;
}

extern "C" int s2m_multi_create(const int* device_ordinals, int n, uint32_t flags, s2m_multi** out) {
  if (!out) return fail(S2M_ERR_INVALID_ARG, "s2m_multi_create: out is NULL");
  *out = nullptr;
  if (!device_ordinals || n < 1 || n > 64) return fail(S2M_ERR_INVALID_ARG, "s2m_multi_create: 1 .. 64 device ordinals");
  // One ordinal may appear several times only when the counts travel through host memory (NCCL refuses two ranks on one
  // device): several slabs then share a GPU, each with its own context, streams and buffers -- how the z-slab path is
  // exercised on a one-GPU box (tests/test_multi_gpu.py).
  for (int a = 0; a < n; ++a)
    for (int b = a + 1; b < n; ++b)
      if (device_ordinals[a] == device_ordinals[b] && !(flags & S2M_MULTI_NO_NCCL))
        return fail(S2M_ERR_INVALID_ARG, "s2m_multi_create: a device ordinal appears twice (allowed only with S2M_MULTI_NO_NCCL)");
  std::unique_ptr<s2m_multi, void (*)(s2m_multi*)> mc(new s2m_multi(), s2m_multi_destroy);
  mc->n = n; mc->flags = flags;
  mc->ordinals.assign(device_ordinals, device_ordinals + n);
  mc->ctx.assign((size_t)n, nullptr);
  mc->comm.assign((size_t)n, nullptr);
  mc->d_mine.assign((size_t)n, nullptr); mc->d_all.assign((size_t)n, nullptr); mc->h_mine.assign((size_t)n, nullptr);
  mc->xs.assign((size_t)n, nullptr);
  mc->counts.assign((size_t)n, 0);
  for (int k = 0; k < n; ++k) {
    mc->workers.emplace_back(new s2m_multi::Worker());
    mc->workers.back()->th = std::thread(worker_loop, mc->workers.back().get());
  }
  // contexts come up concurrently (CUDA context creation is ~100 ms per device)
  std::vector<int> status((size_t)n, S2M_OK);
  std::vector<std::string> errors((size_t)n);
  mc->run_on_all([&](int k) {
    status[(size_t)k] = s2m_ctx_create(mc->ordinals[(size_t)k], &mc->ctx[(size_t)k]);
    if (status[(size_t)k]) { errors[(size_t)k] = s2m_last_error(); return; }
    cudaError_t e = cudaSetDevice(mc->ordinals[(size_t)k]);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&mc->d_mine[(size_t)k]), 8);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&mc->d_all[(size_t)k]), 8 * (size_t)std::max(n, 1));
    if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void**>(&mc->h_mine[(size_t)k]), 8, cudaHostAllocPortable);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&mc->xs[(size_t)k], cudaStreamNonBlocking);
    if (e != cudaSuccess) { status[(size_t)k] = S2M_ERR_CUDA; errors[(size_t)k] = std::string("exchange buffers: ") + cudaGetErrorString(e); }
  });
  for (int k = 0; k < n; ++k) if (status[(size_t)k]) return fail(status[(size_t)k], "device " + std::to_string(mc->ordinals[(size_t)k]) + ": " + errors[(size_t)k]);
  if (n > 1 && !(flags & S2M_MULTI_NO_NCCL)) {
    NcclApi& api = nccl();
    if (!api.ok) return fail(S2M_ERR_UNSUPPORTED, api.why + " (pass S2M_MULTI_NO_NCCL to exchange the counts through host memory)");
    int r = api.CommInitAll(mc->comm.data(), n, mc->ordinals.data());
    if (r != 0) return fail(S2M_ERR_CUDA, std::string("ncclCommInitAll: ") + api.GetErrorString(r));
    api.GetVersion(&mc->nccl_version);
    mc->use_nccl = true;
  }
  *out = mc.release();
  return S2M_OK;
}

extern "C" int s2m_multi_size(const s2m_multi* mc) { return mc ? mc->n : 0; }
extern "C" s2m_ctx* s2m_multi_ctx(s2m_multi* mc, int k) { return (mc && k >= 0 && k < mc->n) ? mc->ctx[(size_t)k] : nullptr; }
extern "C" int s2m_multi_uses_nccl(const s2m_multi* mc, int* nccl_version) {
  if (nccl_version) *nccl_version = mc ? mc->nccl_version : 0;
  return mc && mc->use_nccl ? 1 : 0;
}
extern "C" int s2m_multi_get_partition(const s2m_multi* mc, uint32_t* bounds_out) {
  if (!mc || !bounds_out) return fail(S2M_ERR_INVALID_ARG, "NULL argument");
  if (mc->bounds.size() != (size_t)mc->n + 1) return fail(S2M_ERR_STATE, "no run yet: there is no partition");
  std::copy(mc->bounds.begin(), mc->bounds.end(), bounds_out);
  return S2M_OK;
}
extern "C" int s2m_multi_last_timings(const s2m_multi* mc, s2m_multi_timings* out) {
  if (!mc || !out) return fail(S2M_ERR_INVALID_ARG, "NULL argument");
  *out = mc->last;
  return S2M_OK;
}

extern "C" int s2m_multi_mesh_run(s2m_multi* mc, const s2m_module* compiled, const s2m_mesh_params* p, s2m_result** parts_out) {
  if (!mc || !compiled || !p || !parts_out) return fail(S2M_ERR_INVALID_ARG, "s2m_multi_mesh_run: NULL argument");
  if (p->struct_size != sizeof(s2m_mesh_params)) return fail(S2M_ERR_INVALID_ARG, "s2m_mesh_params.struct_size mismatch");
  if (p->z_begin != 0 || p->z_end != 0) return fail(S2M_ERR_INVALID_ARG, "s2m_multi_mesh_run splits the whole grid itself: z_begin / z_end must be 0");
  const int n = mc->n;
  for (int k = 0; k < n; ++k) parts_out[k] = nullptr;
  const uint32_t n_slices = (p->flags & S2M_MESH_ALL_SLICES) ? p->dims[2] : p->dims[2] - 1u;
  if (p->dims[2] < 2 || n_slices < (uint32_t)n) return fail(S2M_ERR_INVALID_ARG, "fewer z-slices than GPUs");
  std::vector<int> status((size_t)n, S2M_OK);
  std::vector<std::string> errors((size_t)n);
  auto first_error = [&](const char* what) -> int {
    for (int k = 0; k < n; ++k)
      if (status[(size_t)k]) return fail(status[(size_t)k], std::string(what) + ", device " + std::to_string(mc->ordinals[(size_t)k]) + ": " + errors[(size_t)k]);
    return S2M_OK;
  };
  // ---- modules: the compiled cubins loaded once per device
  const uint64_t uid = s2m_module_uid(compiled);
  auto it = mc->modules.find(uid);
  if (it == mc->modules.end()) {
    std::vector<s2m_module*> mods((size_t)n, nullptr);
    mc->run_on_all([&](int k) {
      status[(size_t)k] = s2m_module_instantiate(compiled, mc->ctx[(size_t)k], &mods[(size_t)k]);
      if (status[(size_t)k]) errors[(size_t)k] = s2m_last_error();
    });
    if (int st = first_error("s2m_module_instantiate")) { for (s2m_module* m : mods) if (m) s2m_module_free(m); return st; }
    if (mc->module_order.size() >= 8) {
      const uint64_t old = mc->module_order.front();
      mc->module_order.erase(mc->module_order.begin());
      for (s2m_module* m : mc->modules[old]) if (m) s2m_module_free(m);
      mc->modules.erase(old);
    }
    mc->module_order.push_back(uid);
    it = mc->modules.emplace(uid, std::move(mods)).first;
  }
  std::vector<s2m_module*>& mods = it->second;
  // ---- partition: cost probe on the first GPU when the (module, grid) changes
  if (mc->part_module != uid || !same_grid(mc->part_params, *p) || mc->bounds.size() != (size_t)n + 1) {
    mc->bounds.assign((size_t)n + 1, 0);
    mc->cost.clear();
    if (n > 1 && !(mc->flags & S2M_MULTI_EQUAL_SLABS)) {
      mc->cost.assign(128, 0.0);
      if (s2m_cost_probe(mc->ctx[0], mods[0], p, 128, mc->cost.data()) != S2M_OK) mc->cost.clear();
    }
    int st = s2m_partition_slices(n_slices, n, mc->cost.empty() ? nullptr : mc->cost.data(), (int)mc->cost.size(), mc->bounds.data());
    if (st) return st;
    mc->part_module = uid; mc->part_params = *p; mc->runs_on_partition = 0;
  }
  // ---- the run: begin on every GPU, one all-gather of the vertex counts, finish with the slab's base
  std::vector<double> begin_ms((size_t)n, 0.0), exch_ms((size_t)n, 0.0), finish_ms((size_t)n, 0.0);
  std::vector<s2m_result*> res((size_t)n, nullptr);
  std::atomic<int> failed{0};
  const double t_run0 = now_ms();
  mc->run_on_all([&](int k) {
    const size_t K = (size_t)k;
    s2m_mesh_params q = *p;
    q.z_begin = mc->bounds[K]; q.z_end = mc->bounds[K + 1];
    q.flags |= S2M_MESH_RELATIVE_QUADS;   // nothing waits for the base: s2m_write_mesh_parts and quad_index_add understand it
    const double t0 = now_ms();
    int st = s2m_mesh_begin(mc->ctx[K], mods[K], &q, &res[K]);
    if (st) { status[K] = st; errors[K] = s2m_last_error(); failed.store(1); }
    s2m_result_info info;
    memset(&info, 0, sizeof info);
    if (!st) s2m_result_get(res[K], &info);
    const double t1 = now_ms();
    begin_ms[K] = t1 - t0;
    // every thread must take part in the exchange even after a failure, or the others would wait forever
    uint64_t all[64];
    if (mc->use_nccl) {
      cudaSetDevice(mc->ordinals[K]);
      cudaStream_t s = mc->xs[K];
      *mc->h_mine[K] = st ? ~0ull : info.n_vertices;
      cudaError_t e = cudaMemcpyAsync(mc->d_mine[K], mc->h_mine[K], 8, cudaMemcpyHostToDevice, s);
      int r = nccl().AllGather(mc->d_mine[K], mc->d_all[K], 1, kNcclUint64, mc->comm[K], s);
      if (r != 0 && !status[K]) { status[K] = S2M_ERR_CUDA; errors[K] = std::string("ncclAllGather: ") + nccl().GetErrorString(r); failed.store(1); }
      // the result comes back through a one-warp kernel writing to mapped pinned memory, not a D2H cudaMemcpy: that
      // would queue on the copy engine behind the slab's vertex copies
      if (e == cudaSuccess && r == 0) {
        int rs = s2m_read_device_words(mc->ctx[K], mc->d_all[K], (uint32_t)n, all, s);
        if (rs && !status[K]) { status[K] = rs; errors[K] = s2m_last_error(); failed.store(1); }
      } else if (!status[K]) { status[K] = S2M_ERR_CUDA; errors[K] = std::string("count upload: ") + cudaGetErrorString(e); failed.store(1); }
    } else {
      mc->counts[K] = st ? ~0ull : info.n_vertices;
      mc->barrier.wait((unsigned)n);
      for (int j = 0; j < n; ++j) all[j] = mc->counts[(size_t)j];
      mc->barrier.wait((unsigned)n);   // nobody overwrites counts[] for the next run before everyone has read them
    }
    const double t2 = now_ms();
    exch_ms[K] = t2 - t1;
    bool any_bad = false;
    int64_t base = 0;
    for (int j = 0; j < n; ++j) { if (all[j] == ~0ull) any_bad = true; else if (j < k) base += (int64_t)all[j]; }
    if (!st && !any_bad && !status[K]) {
      st = s2m_mesh_finish(res[K], base);
      if (st) { status[K] = st; errors[K] = s2m_last_error(); failed.store(1); }
    }
    finish_ms[K] = now_ms() - t2;
  });
  const double t_run1 = now_ms();
  if (failed.load()) {
    for (s2m_result* r : res) if (r) s2m_result_free(r);
    if (int st = first_error("s2m_multi_mesh_run")) return st;
    return fail(S2M_ERR_STATE, "s2m_multi_mesh_run: a slab failed");
  }
  for (int k = 0; k < n; ++k) parts_out[k] = res[(size_t)k];
  s2m_multi_timings& t = mc->last;
  memset(&t, 0, sizeof t);
  t.n = n;
  t.wall_ms = t_run1 - t_run0;
  for (int k = 0; k < n && k < 64; ++k) { t.begin_ms[k] = begin_ms[(size_t)k]; t.exchange_ms[k] = exch_ms[(size_t)k]; t.finish_ms[k] = finish_ms[(size_t)k]; }
  // ---- the boundaries are refined from the second and the fourth run on a partition (the first pays for allocations,
  //      the run after a refinement for re-allocations)
  ++mc->runs_on_partition;
  if (n > 1 && (mc->runs_on_partition == 2 || mc->runs_on_partition == 4) && !(mc->flags & (S2M_MULTI_EQUAL_SLABS | S2M_MULTI_NO_REBALANCE))) {
    std::vector<double> sec((size_t)n);
    for (int k = 0; k < n; ++k) sec[(size_t)k] = (begin_ms[(size_t)k] + finish_ms[(size_t)k]) * 1e-3;   // until the slab is complete in host memory
    std::vector<uint32_t> nb((size_t)n + 1);
    if (s2m_rebalance_slices(mc->bounds.data(), n, sec.data(), mc->cost.empty() ? nullptr : mc->cost.data(), (int)mc->cost.size(), nb.data()) == S2M_OK) mc->bounds = nb;
  }
  return S2M_OK;
}
