/* s2m_scan.cuh -- single-pass "decoupled look-back" prefix sum plumbing (Merrill & Garland),
 * shared by the statically compiled kernels (kernels_static.cu) and the NVRTC-compiled ones
 * (kernels_jit.cuh).  Replaces the ordering work the reference does on the host:
 *   /root/reference/src/mesh.rs:237-245  VertexList::insert (append in scan order)
 *   /root/reference/src/mesh.rs:327-331  vertex_index (binary search over that order)
 *
 * One 64-bit status word per tile: bits 63..62 = state (0 empty, 1 aggregate, 2 inclusive prefix),
 * bits 61..0 = value.  A single 64-bit relaxed load/store carries state and value together, so no
 * fence is needed between them.  Tiles take their index from an atomic ticket so a tile's
 * predecessors are always already running (forward progress without co-residency assumptions).
 */
#ifndef S2M_SCAN_CUH_
#define S2M_SCAN_CUH_

#define S2M_SCAN_EMPTY 0ull
#define S2M_SCAN_AGGREGATE 1ull
#define S2M_SCAN_PREFIX 2ull
#define S2M_SCAN_VALUE_MASK 0x3fffffffffffffffull

__device__ __forceinline__ unsigned long long s2m_ld_relaxed(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void s2m_st_relaxed(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

/* The look-back in two halves, for kernels that know their tile's total long before they need its prefix (K4a: the
 * vertex count is known after the corner evaluations, the prefix is needed after the normal taps): publish early, look
 * back late -- by then the predecessors have published too and nobody spins.
 *
 * s2m_publish_aggregate: lane 0 of warp 0 (any one thread) announces the tile's total. */
__device__ __forceinline__ void s2m_publish_aggregate(unsigned long long* status, unsigned tile, unsigned long long aggregate,
                                                      unsigned long long base) {
  if (tile == 0) s2m_st_relaxed(status, (S2M_SCAN_PREFIX << 62) | (base + aggregate));
  else s2m_st_relaxed(status + tile, (S2M_SCAN_AGGREGATE << 62) | aggregate);
}
/* s2m_lookback_published: called by every thread of warp 0 of the block (all 32 lanes converged) after
 * s2m_publish_aggregate.  `aggregate` is the block's total (same value in all lanes).  Returns the exclusive prefix of
 * this tile, i.e. `base` plus the aggregates of all earlier tiles, and upgrades the tile's status to that prefix. */
__device__ __forceinline__ unsigned long long s2m_lookback_published(unsigned long long* status, unsigned tile,
                                                                     unsigned long long aggregate,
                                                                     unsigned long long base) {
  const unsigned lane = threadIdx.x & 31u;
  if (tile == 0) return base;
  unsigned long long exclusive = 0;
  int look = (int)tile - 1;  /* nearest predecessor handled by lane 0 */
  for (;;) {
    int t = look - (int)lane;
    unsigned long long s = (S2M_SCAN_PREFIX << 62); /* lanes before tile 0: prefix 0 ... */
    if (t >= 0) {
      do { s = s2m_ld_relaxed(status + t); } while ((s >> 62) == S2M_SCAN_EMPTY);
    } else if (t < -1) {
      s = (S2M_SCAN_AGGREGATE << 62); /* ... but only the first out-of-range lane stops the walk */
    }
    /* t == -1 plays the role of a virtual tile holding prefix `base` */
    unsigned long long val = (t >= 0) ? (s & S2M_SCAN_VALUE_MASK) : (t == -1 ? base : 0ull);
    unsigned has_prefix = __ballot_sync(0xffffffffu, (s >> 62) == S2M_SCAN_PREFIX);
    unsigned first = has_prefix ? (unsigned)(__ffs((int)has_prefix) - 1) : 32u; /* nearest lane with a prefix */
    unsigned long long contrib = (lane <= first) ? val : 0ull;
    for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
    exclusive += contrib;
    if (has_prefix) break;
    look -= 32;
  }
  if (lane == 0) s2m_st_relaxed(status + tile, (S2M_SCAN_PREFIX << 62) | (exclusive + aggregate));
  return exclusive;
}
/* Both halves back to back.  Called by every thread of warp 0 of the block (all 32 lanes converged). */
__device__ __forceinline__ unsigned long long s2m_lookback_warp(unsigned long long* status, unsigned tile,
                                                                unsigned long long aggregate,
                                                                unsigned long long base) {
  if ((threadIdx.x & 31u) == 0) s2m_publish_aggregate(status, tile, aggregate, base);
  return s2m_lookback_published(status, tile, aggregate, base);
}

/* Block-wide exclusive scan of one unsigned per thread (blockDim.x multiple of 32, <= 1024).
 * smem: at least 33 unsigned.  Returns the exclusive prefix; *total gets the block sum. */
__device__ __forceinline__ unsigned s2m_block_exclusive_scan(unsigned v, unsigned* smem, unsigned* total) {
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31u) >> 5;
  unsigned inc = v;
  for (int o = 1; o < 32; o <<= 1) {
    unsigned n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= (unsigned)o) inc += n;
  }
  if (lane == 31) smem[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    unsigned w = (lane < nwarps) ? smem[lane] : 0u;
    unsigned winc = w;
    for (int o = 1; o < 32; o <<= 1) {
      unsigned n = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= (unsigned)o) winc += n;
    }
    smem[lane] = winc - w; /* exclusive warp offsets */
    if (lane == 31) smem[32] = winc;
  }
  __syncthreads();
  unsigned excl = smem[warp] + inc - v;
  *total = smem[32];
  return excl;
}

#endif /* S2M_SCAN_CUH_ */
