// kernels_static.cu -- the SDF-independent kernels, compiled ahead of time for sm_100a.
//
// K2  k2_classify   replaces the per-cell sign tests + host pixel scan that decide where the
//                   surface is (/root/reference/src/bin/sdf2mesh/dualcontour.wgsl:57-69, :119-128;
//                   /root/reference/src/bin/sdf2mesh/main.rs:327-344) by a conservative CANDIDATE
//                   test on the once-per-corner slab: a cell is a candidate unless all 8 corners
//                   are > +tau or all 8 are < -tau.  Exact reference arithmetic is then re-run on
//                   candidates only (K4a).  Corner classes are formed once per corner, packed with
//                   __ballot_sync into row bitmasks staged in shared memory, and combined per cell
//                   with word-wide AND/shift; counts with __popc.
// K3  k3_compact    single-pass decoupled look-back exclusive scan of the candidate bitmask:
//                   writes the per-word rank table and the compact candidate list in linear cell
//                   order == the reference's VertexList order (mesh.rs:224-245).
// K4b k4_quads      replaces VertexList::fetch_triangle_indices + vertex_index + Quad::swap /
//                   is_valid (/root/reference/src/mesh.rs:267-331, /root/reference/src/lib.rs:187-211):
//                   3 edge tests per vertex, neighbour ranks by bitmask rank lookup instead of binary
//                   search, stable single-pass emission of 64-bit-index quads.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "kernels_static.h"
#include "s2m_scan.cuh"

namespace {

constexpr int K2_TX_WORDS = 4;   // 128 cells in x per CTA
constexpr int K2_TY = 16;        // cell rows per CTA
constexpr int K2_ZT = 16;        // cell slices marched per CTA

// End of a K2 block (one thread): add the block's candidate count; the block that finishes LAST copies the total into
// mapped pinned host memory, so the host reads it after the stream's event without a launch or a cudaMemcpy of its own
// (a D2H memcpy of 8 bytes would queue on the copy engine behind the previous chunk's bulk copy).
__device__ __forceinline__ void k2_finish(unsigned block_count, unsigned long long* total, unsigned* done, unsigned long long* host_total) {
  if (block_count) atomicAdd(total, (unsigned long long)block_count);
  if (done == nullptr) return;
  __threadfence();
  const unsigned n_blocks = gridDim.x * gridDim.y * gridDim.z;
  if (atomicAdd(done, 1u) == n_blocks - 1u) {
    __threadfence();
    const unsigned long long t = atomicAdd(total, 0ull);
    *host_total = t;
    __threadfence_system();
  }
}

__device__ __forceinline__ uint32_t pair_x(const uint32_t* row, int i) {
  // bit b of the result: corner (32*i + b) AND corner (32*i + b + 1)
  return row[i] & ((row[i] >> 1) | (row[i + 1] << 31));
}

__global__ void __launch_bounds__(256)
k2_classify(const float* __restrict__ slab, uint32_t pitch_x, unsigned long long plane_stride,
            uint32_t res_x, uint32_t res_y, uint32_t nz_chunk, float tau,
            uint32_t* __restrict__ cand_mask, uint32_t words_x, unsigned long long* __restrict__ total,
            unsigned* __restrict__ done, unsigned long long* __restrict__ host_total) {
  __shared__ uint32_t sP[2][K2_TY + 1][K2_TX_WORDS + 1];
  __shared__ uint32_t sN[2][K2_TY + 1][K2_TX_WORDS + 1];
  __shared__ unsigned s_red[8];
  const uint32_t x0 = blockIdx.x * (32u * K2_TX_WORDS);
  const uint32_t y0 = blockIdx.y * K2_TY;
  const uint32_t zt0 = blockIdx.z * K2_ZT;
  const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  unsigned count = 0;

  for (uint32_t k = 0; k <= (uint32_t)K2_ZT; ++k) {
    const uint32_t plane = zt0 + k;
    if (plane > nz_chunk) break;  // uniform across the block
    const int buf = (int)(k & 1u);
    const float* pl = slab + (unsigned long long)plane * plane_stride;
    for (uint32_t row = warp; row <= (uint32_t)K2_TY; row += 8u) {
      const uint32_t y = y0 + row;
      const bool yok = y <= res_y;
      const float* rp = pl + (unsigned long long)y * pitch_x;
      float v[K2_TX_WORDS];
#pragma unroll
      for (int i = 0; i < K2_TX_WORDS; ++i) {
        const uint32_t x = x0 + 32u * i + lane;
        v[i] = (yok && x <= res_x) ? __ldg(rp + x) : __int_as_float(0x7fc00000);
      }
      const uint32_t xh = x0 + 32u * K2_TX_WORDS;
      float vh = __int_as_float(0x7fc00000);
      if (lane == 0 && yok && xh <= res_x) vh = __ldg(rp + xh);
#pragma unroll
      for (int i = 0; i < K2_TX_WORDS; ++i) {
        const uint32_t pw = __ballot_sync(0xffffffffu, v[i] > tau);
        const uint32_t nw = __ballot_sync(0xffffffffu, v[i] < -tau);
        if (lane == 0) { sP[buf][row][i] = pw; sN[buf][row][i] = nw; }
      }
      if (lane == 0) { sP[buf][row][K2_TX_WORDS] = vh > tau ? 1u : 0u; sN[buf][row][K2_TX_WORDS] = vh < -tau ? 1u : 0u; }
    }
    __syncthreads();
    if (k >= 1) {
      const uint32_t z = zt0 + k - 1;  // cell slice inside the chunk (z < nz_chunk because plane <= nz_chunk)
      for (uint32_t idx = threadIdx.x; idx < (uint32_t)(K2_TY * K2_TX_WORDS); idx += blockDim.x) {
        const uint32_t row = idx / K2_TX_WORDS;
        const int i = (int)(idx % K2_TX_WORDS);
        const uint32_t y = y0 + row;
        const uint32_t xw = x0 / 32u + (uint32_t)i;
        if (y < res_y && xw < words_x) {
          const uint32_t allp = pair_x(sP[buf ^ 1][row], i) & pair_x(sP[buf ^ 1][row + 1], i) &
                                pair_x(sP[buf][row], i) & pair_x(sP[buf][row + 1], i);
          const uint32_t alln = pair_x(sN[buf ^ 1][row], i) & pair_x(sN[buf ^ 1][row + 1], i) &
                                pair_x(sN[buf][row], i) & pair_x(sN[buf][row + 1], i);
          uint32_t cand = ~(allp | alln);
          const uint32_t xb = xw * 32u;
          if (xb + 32u > res_x) cand &= (1u << (res_x - xb)) - 1u;  // cells beyond the grid
          cand_mask[((unsigned long long)z * res_y + y) * words_x + xw] = cand;
          count += (unsigned)__popc(cand);
        }
      }
    }
    __syncthreads();
  }
  for (int o = 16; o > 0; o >>= 1) count += __shfl_xor_sync(0xffffffffu, count, o);
  if (lane == 0) s_red[warp] = count;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned t = 0;
    for (int w = 0; w < 8; ++w) t += s_red[w];
    k2_finish(t, total, done, host_total);
  }
}

// K2 (default): the same candidate test on the corner-class planes K1 wrote next to the slab.
// Layout: one 64-bit word per 32 corners of a row; byte l holds corners 4l..4l+3, low nibble
// P (v > tau), high nibble N (v < -tau).  P and N are processed together in 64-bit operations.
// One thread owns one 32-cell word of one cell row and marches through the chunk's slices, keeping
// the previous plane's "both rows, both x-neighbours" word in registers.  0.25 B/voxel read
// instead of 4.
__device__ __forceinline__ unsigned long long cls_pair_x(unsigned long long v, unsigned long long vnext) {
  // class of the NEXT corner in x moved onto each corner's position, then ANDed with its own
  const unsigned long long s = ((v >> 1) & 0x7777777777777777ull) | ((v >> 5) & 0x8888888888888888ull) | ((vnext & 0x11ull) << 59);
  return v & s;
}
__device__ __forceinline__ unsigned long long ld_u64(const uint2* p) {
  const uint2 t = __ldg(p);
  return (unsigned long long)t.x | ((unsigned long long)t.y << 32);
}

// a cell is NOT a candidate iff its P bit or its N bit survived in all 8 corners; the nibble per byte that is left is
// squeezed to 32 bits with one OR of the word shifted by a nibble and a byte permute (three shift / or / and rounds on
// 64 bits before: K2 is issue-bound, profiles/r02_ncu_k234.md)
__device__ __forceinline__ uint32_t cls_candidates(unsigned long long prev, unsigned long long cur) {
  const unsigned long long all8 = prev & cur;                       // low nibbles: all 8 corners > tau; high: all < -tau
  const unsigned long long x = ~(all8 | (all8 >> 4)) & 0x0f0f0f0f0f0f0f0full;   // nibble per byte: candidate cells
  const unsigned long long y = x | (x >> 4);                        // bytes 0, 2, 4, 6 now hold cells 0-7, 8-15, 16-23, 24-31
  return __byte_perm((uint32_t)y, (uint32_t)(y >> 32), 0x6420);
}

__global__ void __launch_bounds__(256)
k2_classify_bits(const uint2* __restrict__ cls, uint32_t cls_words, uint32_t rows, uint32_t res_x, uint32_t res_y,
                 uint32_t nz_chunk, uint32_t* __restrict__ cand_mask, uint32_t words_x, unsigned long long* __restrict__ total,
                 unsigned* __restrict__ done, unsigned long long* __restrict__ host_total, uint32_t* __restrict__ seg_count) {
  __shared__ unsigned s_red[8];
  const uint32_t xw = blockIdx.x * 32u + (threadIdx.x & 31u);
  const uint32_t y = blockIdx.y * 8u + (threadIdx.x >> 5);
  const uint32_t zt0 = blockIdx.z * K2_ZT;
  const bool live = xw < words_x && y < res_y;
  const bool has_next = xw + 1u < cls_words;
  unsigned count = 0;
  unsigned long long prev = 0;
  // the thread's words of plane zt0 and where slice zt0's results go; every plane / slice further is one stride on
  // (the addresses were a quarter of K2's instructions when they were computed from the indices in every iteration)
  const uint2* r0 = cls + ((unsigned long long)zt0 * rows + y) * cls_words + xw;
  const unsigned long long plane_step = (unsigned long long)rows * cls_words;
  uint32_t* out = cand_mask + ((unsigned long long)zt0 * res_y + y) * words_x + xw;
  const unsigned long long out_step = (unsigned long long)res_y * words_x;
  uint32_t* seg = seg_count != nullptr ? seg_count + ((unsigned long long)zt0 * res_y + y) * gridDim.x + blockIdx.x : nullptr;
  const unsigned long long seg_step = (unsigned long long)res_y * gridDim.x;
  const uint32_t xb = xw * 32u;
  const uint32_t edge = xb + 32u > res_x ? (xb < res_x ? (1u << (res_x - xb)) - 1u : 0u) : 0xffffffffu;   // cells beyond the grid
  const uint32_t n_planes = min((uint32_t)K2_ZT, nz_chunk - min(zt0, nz_chunk)) + 1u;   // planes zt0 .. zt0 + n_planes - 1 <= nz_chunk
  for (uint32_t k = 0; k < n_planes; ++k) {
    unsigned long long cur = 0;
    unsigned pc = 0;
    if (live) {
      const uint2* r1 = r0 + cls_words;
      const unsigned long long a = ld_u64(r0), c = ld_u64(r1);
      const unsigned long long b = has_next ? ld_u64(r0 + 1) : 0ull, d = has_next ? ld_u64(r1 + 1) : 0ull;
      // P nibbles: all 4 corners of this plane > tau; N nibbles: all < -tau.  AND the two corner rows first, then pair in
      // x once: the shift inside the pairing distributes over AND, so pair(a, b) & pair(c, d) == pair(a & c, b & d).
      cur = cls_pair_x(a & c, b & d);
      if (k >= 1) {
        const uint32_t cand = cls_candidates(prev, cur) & edge;
        *out = cand;
        pc = (unsigned)__popc(cand);
        count += pc;
      }
    }
    // a warp is one segment (32 consecutive words of one cell row): its candidate count, for K3's scan
    if (k >= 1) {
      if (seg != nullptr) {
        const unsigned sc = __reduce_add_sync(0xffffffffu, pc);
        if ((threadIdx.x & 31u) == 0 && y < res_y) *seg = sc;
        seg += seg_step;
      }
      out += out_step;
    }
    r0 += plane_step;
    prev = cur;
  }
  for (int o = 16; o > 0; o >>= 1) count += __shfl_xor_sync(0xffffffffu, count, o);
  if ((threadIdx.x & 31u) == 0) s_red[threadIdx.x >> 5] = count;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned t = 0;
    for (int w = 0; w < 8; ++w) t += s_red[w];
    k2_finish(t, total, done, host_total);
  }
}

// ------------------------------------------------------------------------------------------ K3
// Compaction of the candidate bits into the ordered candidate list + the rank table, in ONE pass over per-segment
// counts instead of the mask.  A segment is 32 consecutive mask words of one cell row (1024 cells, one 128-byte line);
// K2 writes every segment's candidate count while it writes the mask (k_seg_count does it for the other mask sources).
// Segments in index order are mask words in index order, so: block scan of the counts + decoupled look-back across
// tiles gives every segment its first rank; then the tile's NON-EMPTY segments -- at 2048^3 the surface passes through
// a fraction of them -- are expanded by the block's warps, one segment per warp at a time: lane l owns word l, a warp
// scan of the popcounts gives the word's rank (word_prefix) and each lane appends its cells' keys.  The mask is read
// only where it has bits (the first version scanned all of it: 1.3 ms at 2048^3, 4 barriers per 16 KB of mask).
constexpr int K3_THREADS = 256;
constexpr int K3_SPT = 2;                         // segments per thread
constexpr int K3_TILE_SEGS = K3_THREADS * K3_SPT; // 512 segments = 16384 mask words per tile

__global__ void __launch_bounds__(32 * 8)
k_seg_count(const uint32_t* __restrict__ cand_mask, unsigned long long n_rows, uint32_t words_x, uint32_t segs_x, uint32_t* __restrict__ seg_count) {
  const unsigned long long seg = (unsigned long long)blockIdx.x * 8u + (threadIdx.x >> 5);
  if (seg >= n_rows * segs_x) return;
  const unsigned long long row = seg / segs_x;
  const uint32_t xw = (uint32_t)(seg - row * segs_x) * 32u + (threadIdx.x & 31u);
  const unsigned pc = xw < words_x ? (unsigned)__popc(cand_mask[row * words_x + xw]) : 0u;
  const unsigned sc = __reduce_add_sync(0xffffffffu, pc);
  if ((threadIdx.x & 31u) == 0) seg_count[seg] = sc;
}

__global__ void __launch_bounds__(K3_THREADS)
k3_compact(const uint32_t* __restrict__ cand_mask, const uint32_t* __restrict__ seg_count, unsigned long long n_segs, uint32_t words_x,
           uint32_t segs_x, uint32_t res_y, uint32_t z_offset, uint32_t* __restrict__ word_prefix, unsigned long long* __restrict__ cand_key,
           unsigned long long base, unsigned long long* status, unsigned* ticket) {
  __shared__ unsigned s_scan[33];
  __shared__ unsigned s_tile, s_n;
  __shared__ unsigned long long s_base;
  __shared__ unsigned s_seg[K3_TILE_SEGS];   // non-empty segments of the tile (index within the tile) ...
  __shared__ unsigned s_rank[K3_TILE_SEGS];  // ... and the rank of their first candidate within the tile
  if (threadIdx.x == 0) { s_tile = atomicAdd(ticket, 1u); s_n = 0u; }
  __syncthreads();
  const unsigned tile = s_tile;
  const unsigned long long s0 = (unsigned long long)tile * K3_TILE_SEGS + (unsigned long long)threadIdx.x * K3_SPT;
  unsigned c[K3_SPT];
  if (s0 + K3_SPT <= n_segs) {   // seg_count regions start at multiples of the per-slice segment count: 8-byte aligned when that is even
    if ((reinterpret_cast<unsigned long long>(seg_count) & 7ull) == 0ull) {
      const uint2 t = __ldg(reinterpret_cast<const uint2*>(seg_count + s0));
      c[0] = t.x; c[1] = t.y;
    } else {
#pragma unroll
      for (int q = 0; q < K3_SPT; ++q) c[q] = __ldg(seg_count + s0 + q);
    }
  } else {
#pragma unroll
    for (int q = 0; q < K3_SPT; ++q) c[q] = (s0 + q < n_segs) ? __ldg(seg_count + s0 + q) : 0u;
  }
  static_assert(K3_SPT == 2, "the vector load above reads 2 counts");
  unsigned mine = 0;
#pragma unroll
  for (int q = 0; q < K3_SPT; ++q) mine += c[q];
  unsigned total = 0;
  unsigned local = s2m_block_exclusive_scan(mine, s_scan, &total);
  if (threadIdx.x < 32) {
    unsigned long long b = s2m_lookback_warp(status, tile, (unsigned long long)total, base);
    if (threadIdx.x == 0) s_base = b;
  }
  if (total == 0u) return;   // uniform over the block: nothing to expand in this tile
#pragma unroll
  for (int q = 0; q < K3_SPT; ++q) {
    if (c[q]) {
      const unsigned at = atomicAdd(&s_n, 1u);
      s_seg[at] = threadIdx.x * K3_SPT + (unsigned)q;
      s_rank[at] = local;
    }
    local += c[q];
  }
  __syncthreads();
  const unsigned n = s_n;
  const unsigned long long tile_base = s_base;
  const unsigned lane = threadIdx.x & 31u;
  // K3_BATCH segments per warp and round: their mask lines are requested together (a warp that expands one segment at
  // a time has one 128-byte load in flight and waits a full memory latency per segment)
  constexpr unsigned K3_BATCH = 4;
  for (unsigned i0 = threadIdx.x >> 5; i0 < n; i0 += (K3_THREADS / 32) * K3_BATCH) {
    unsigned long long rowi[K3_BATCH], w[K3_BATCH];
    uint32_t xw[K3_BATCH], bits[K3_BATCH];
    unsigned rank0[K3_BATCH];
#pragma unroll
    for (unsigned j = 0; j < K3_BATCH; ++j) {
      const unsigned i = i0 + j * (K3_THREADS / 32);
      bits[j] = 0u; rowi[j] = 0ull; w[j] = 0ull; xw[j] = 0xffffffffu; rank0[j] = 0u;
      if (i < n) {
        const unsigned long long seg = (unsigned long long)tile * K3_TILE_SEGS + s_seg[i];
        rowi[j] = seg / segs_x;
        xw[j] = (uint32_t)(seg - rowi[j] * segs_x) * 32u + lane;
        w[j] = rowi[j] * words_x + xw[j];
        rank0[j] = s_rank[i];
        if (xw[j] < words_x) bits[j] = __ldg(cand_mask + w[j]); else xw[j] = 0xffffffffu;
      }
    }
#pragma unroll
    for (unsigned j = 0; j < K3_BATCH; ++j) {
      if (i0 + j * (K3_THREADS / 32) >= n) break;   // uniform over the warp
      uint32_t b = bits[j];
      const unsigned pc = (unsigned)__popc(b);
      unsigned inc = pc;
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (unsigned)o) inc += t;
      }
      unsigned long long run = tile_base + rank0[j] + (inc - pc);
      if (xw[j] != 0xffffffffu) word_prefix[w[j]] = (uint32_t)run;   // the whole line: rank_of (K4b) reads it only for words with a set bit
      if (b) {
        const uint32_t z = (uint32_t)(rowi[j] / res_y);
        const uint32_t y = (uint32_t)(rowi[j] - (unsigned long long)z * res_y);
        const unsigned long long hi = ((unsigned long long)y << 16) | ((unsigned long long)(z + z_offset) << 32);
        while (b) {
          const int bit = __ffs((int)b) - 1;
          b &= b - 1u;
          cand_key[run++] = hi | (unsigned long long)(xw[j] * 32u + (uint32_t)bit);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ K4b
struct QuadParams {
  const unsigned long long* vert_key;   // label keys
  const unsigned char* vert_nibble;
  // Ranges live in DEVICE memory, so that the launch needs no host round-trip after K4a: tot_prev / tot_cur point at
  // {vertices, quads} emitted up to the end of the previous / this z-chunk (tot_cur[0] was written by K4a, tot_cur[1]
  // is written here), n_halo at the number of vertices of the recomputed slice below the slab (they come first).
  const unsigned long long* tot_prev;
  unsigned long long* tot_cur;
  const unsigned long long* n_halo;
  const uint32_t* cand_mask;
  const uint32_t* word_prefix;
  const uint32_t* cand_vrank;
  uint32_t words_x, res_y;
  uint32_t z_first;       // true z of the first slice present in cand_mask
  uint32_t label_add;     // label = true z + label_add
  long long index_add;    // emitted index = local vertex index - n_halo + index_add  (slab-relative when 0)
  unsigned long long* quads;  // 4 per quad
  uint32_t* quads32;          // S2M_MESH_QUADS_U32: 4 x u32 per quad instead (the reference's index type, lib.rs Quad)
  unsigned long long* status;
  unsigned* ticket;
  unsigned long long* n_invalid;
  unsigned long long* invalid_records;
  unsigned long long* invalid_cursor;
  unsigned long long invalid_capacity;
  unsigned* done;                 // optional: zeroed block-completion counter
  unsigned long long* host_slot;  // mapped pinned host memory, 5 words (k4b_finish)
};

constexpr uint32_t MISSING32 = 0xffffffffu;

__device__ __forceinline__ uint32_t rank_of(const QuadParams& p, int x, int y, int z_true) {
  // vertex index (local) of cell (x, y, z_true) or MISSING32.  z_true may be z_first-1 -> missing.
  if (x < 0 || y < 0 || z_true < (int)p.z_first) return MISSING32;
  const unsigned long long w = ((unsigned long long)(z_true - (int)p.z_first) * p.res_y + (unsigned)y) * p.words_x + ((unsigned)x >> 5);
  const uint32_t m = __ldg(p.cand_mask + w);
  const uint32_t bit = 1u << (x & 31);
  if (!(m & bit)) return MISSING32;
  const uint32_t c = __ldg(p.word_prefix + w) + (uint32_t)__popc(m & (bit - 1u));
  return __ldg(p.cand_vrank + c);
}

// End of a K4b block (one thread).  The block that finishes LAST publishes the chunk's totals into mapped pinned host
// memory: {vertices, quads} up to the end of this chunk, halo vertices, invalid quads, invalid records -- what the host
// needs to size the chunk's device->host copies, without a launch or a cudaMemcpy of its own.
__device__ __forceinline__ void k4b_finish(const QuadParams& p) {
  if (p.done == nullptr) return;
  __threadfence();
  if (atomicAdd(p.done, 1u) == gridDim.x - 1u) {
    __threadfence();
    p.host_slot[0] = s2m_ld_relaxed(p.tot_cur);
    p.host_slot[1] = s2m_ld_relaxed(p.tot_cur + 1);
    p.host_slot[2] = s2m_ld_relaxed(p.n_halo);
    p.host_slot[3] = s2m_ld_relaxed(p.n_invalid);
    p.host_slot[4] = p.invalid_cursor ? s2m_ld_relaxed(p.invalid_cursor) : 0ull;
    __threadfence_system();
  }
}

__global__ void __launch_bounds__(256)
k4_quads(QuadParams p) {
  __shared__ unsigned s_scan[33];
  __shared__ unsigned s_tile;
  __shared__ unsigned long long s_base;
  __shared__ unsigned s_inv[8];
  if (threadIdx.x == 0) s_tile = atomicAdd(p.ticket, 1u);
  __syncthreads();
  const unsigned tile = s_tile;
  const unsigned long long n_halo = *p.n_halo;
  const unsigned long long v_end = p.tot_cur[0];
  const unsigned long long v_begin = p.tot_prev[0] > n_halo ? p.tot_prev[0] : n_halo;
  const unsigned long long quad_base = p.tot_prev[1];
  const unsigned long long n_own = v_end > v_begin ? v_end - v_begin : 0ull;
  const long long index_offset = p.index_add - (long long)n_halo;
  // The grid is sized from an upper bound (the chunk's candidate count): tiles past the vertices have nothing to
  // do and nobody looks back at them.  With no vertex at all, tile 0 carries the quad total forward.
  if ((unsigned long long)tile * blockDim.x >= n_own) {
    if (tile == 0 && threadIdx.x == 0) p.tot_cur[1] = quad_base;
    if (threadIdx.x == 0) k4b_finish(p);
    return;
  }
  const unsigned long long i = (unsigned long long)tile * blockDim.x + threadIdx.x;  // vertex ordinal in this launch
  uint32_t q[3][4];
  unsigned nvalid = 0, ninvalid = 0;
  if (i < n_own) {
    const unsigned long long vi = i + v_begin;
    const unsigned long long key = p.vert_key[vi];
    const int x = (int)(key & 0xffffu), y = (int)((key >> 16) & 0xffffu);
    const uint32_t label = (uint32_t)(key >> 32);
    const int z = (int)(label - p.label_add);  // true z
    const unsigned nib = p.vert_nibble[vi];
    const bool s100 = nib & 1u, s010 = nib & 2u, s001 = nib & 4u, s000 = nib & 8u;
    const uint32_t self = (uint32_t)vi;
    // mesh.rs:286-296  X-edge quad.  The z guard is on the LABEL; a slice below the
    // first scanned one does not exist in the list -> MISSING -> invalid quad, as in the reference.
    auto emit = [&](uint32_t a, uint32_t b, uint32_t c, uint32_t d, bool swap, unsigned edge) {
      if (a == MISSING32 || b == MISSING32 || c == MISSING32 || d == MISSING32) {
        ++ninvalid;
        if (p.invalid_records) {  // rare: a handful per million vertices
          const unsigned long long slot = atomicAdd(p.invalid_cursor, 1ull);
          if (slot < p.invalid_capacity) {
            uint32_t o[4] = {a, b, c, d};
            if (swap) { o[0] = d; o[1] = c; o[2] = b; o[3] = a; }
            unsigned long long* rec = p.invalid_records + 6ull * slot;
            rec[0] = key; rec[1] = edge;
            for (int t = 0; t < 4; ++t) rec[2 + t] = o[t] == MISSING32 ? ~0ull : (unsigned long long)((long long)o[t] + index_offset);
          }
        }
        return;
      }
      uint32_t* o = q[nvalid++];
      if (swap) { o[0] = d; o[1] = c; o[2] = b; o[3] = a; } else { o[0] = a; o[1] = b; o[2] = c; o[3] = d; }
    };
    const bool have_below = (z - 1) >= 0;  // a true slice z-1 exists in the grid
    if (s100 != s000 && y > 0 && label > 0)
      emit(have_below ? rank_of(p, x, y - 1, z - 1) : MISSING32, have_below ? rank_of(p, x, y, z - 1) : MISSING32,
           self, rank_of(p, x, y - 1, z), s100, 0u);
    if (s010 != s000 && x > 0 && label > 0)
      emit(have_below ? rank_of(p, x - 1, y, z - 1) : MISSING32, have_below ? rank_of(p, x, y, z - 1) : MISSING32,
           self, rank_of(p, x - 1, y, z), !s010, 1u);
    if (s001 != s000 && x > 0 && y > 0)
      emit(rank_of(p, x - 1, y - 1, z), rank_of(p, x, y - 1, z), self, rank_of(p, x - 1, y, z), s001, 2u);
  }
  unsigned total = 0;
  const unsigned local = s2m_block_exclusive_scan(nvalid, s_scan, &total);
  if (threadIdx.x < 32) {
    unsigned long long b = s2m_lookback_warp(p.status, tile, (unsigned long long)total, quad_base);
    if (threadIdx.x == 0) s_base = b;
  }
  unsigned inv = ninvalid;
  for (int o = 16; o > 0; o >>= 1) inv += __shfl_xor_sync(0xffffffffu, inv, o);
  if ((threadIdx.x & 31u) == 0) s_inv[threadIdx.x >> 5] = inv;
  __syncthreads();
  unsigned long long at = s_base + local;
  if (p.quads32) {
    for (unsigned k = 0; k < nvalid; ++k, ++at)
      *reinterpret_cast<uint4*>(p.quads32 + 4ull * at) =
          make_uint4((uint32_t)((long long)q[k][0] + index_offset), (uint32_t)((long long)q[k][1] + index_offset),
                     (uint32_t)((long long)q[k][2] + index_offset), (uint32_t)((long long)q[k][3] + index_offset));
  } else
  for (unsigned k = 0; k < nvalid; ++k, ++at) {
    ulonglong2* dst = reinterpret_cast<ulonglong2*>(p.quads + 4ull * at);
    dst[0] = make_ulonglong2((unsigned long long)((long long)q[k][0] + index_offset), (unsigned long long)((long long)q[k][1] + index_offset));
    dst[1] = make_ulonglong2((unsigned long long)((long long)q[k][2] + index_offset), (unsigned long long)((long long)q[k][3] + index_offset));
  }
  if (threadIdx.x == 0) {
    unsigned t = 0;
    for (int w = 0; w < 8; ++w) t += s_inv[w];
    if (t) atomicAdd(p.n_invalid, (unsigned long long)t);
    if ((unsigned long long)(tile + 1) * blockDim.x >= n_own) p.tot_cur[1] = s_base + total;
    k4b_finish(p);
  }
}

// Copies a few counters into mapped pinned host memory with plain stores, so that the host can read
// them after an event WITHOUT a cudaMemcpy: a D2H memcpy would queue on the copy engine behind the
// bulk vertex/quad copies of the previous chunk and stall the pipeline for milliseconds.
__global__ void k_publish(const unsigned long long* __restrict__ src, unsigned long long* __restrict__ dst_host, unsigned n) {
  if (threadIdx.x < n) dst_host[threadIdx.x] = src[threadIdx.x];
  __threadfence_system();
}

// the same from two places: dst = src_a[0..na) ++ src_b[0..nb)
__global__ void k_publish2(const unsigned long long* __restrict__ src_a, unsigned na, const unsigned long long* __restrict__ src_b, unsigned nb,
                           unsigned long long* __restrict__ dst_host) {
  if (threadIdx.x < na) dst_host[threadIdx.x] = src_a[threadIdx.x];
  else if (threadIdx.x < na + nb) dst_host[threadIdx.x] = src_b[threadIdx.x - na];
  __threadfence_system();
}

// FP32 throughput probe (s2m_measure_fp32_peak): CHAINS independent dependent-FMA chains per thread, no memory traffic.
//   MODE 0  FFMA reg,reg,reg        (three register operands: half rate on sm_100, B300_MICROARCH.md "Pipe rates")
//   MODE 1  FFMA reg,imm,imm        (full rate)
//   MODE 2  FFMA2 pair,bcast,imm    (packed f32x2, what K1's packed polynomials issue)
constexpr int FP32_PROBE_CHAINS = 8;
template <int MODE>
__global__ void __launch_bounds__(256) k_fp32_probe(float* out, float b, float c, int iters) {
  float a[FP32_PROBE_CHAINS];
  float2 pr[FP32_PROBE_CHAINS / 2];
  for (int i = 0; i < FP32_PROBE_CHAINS; ++i) a[i] = threadIdx.x * 1e-3f + i;
  for (int i = 0; i < FP32_PROBE_CHAINS / 2; ++i) pr[i] = make_float2(a[2 * i], a[2 * i + 1]);
  const float2 bb = make_float2(b, b);
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < FP32_PROBE_CHAINS; ++i) a[i] = fmaf(a[i], b, c);
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < FP32_PROBE_CHAINS; ++i) a[i] = fmaf(a[i], 0.999f, 1e-3f);
    } else {
#pragma unroll
      for (int i = 0; i < FP32_PROBE_CHAINS / 2; ++i) pr[i] = __ffma2_rn(pr[i], bb, make_float2(1e-3f, 1e-3f));
    }
  }
  float s = 0;
  for (int i = 0; i < FP32_PROBE_CHAINS; ++i) s += a[i];
  for (int i = 0; i < FP32_PROBE_CHAINS / 2; ++i) s += pr[i].x + pr[i].y;
  if (s == 123.456f) out[0] = s;
}

}  // namespace

// ------------------------------------------------------------------------------------------ launchers
extern "C" int s2m_launch_publish2(const unsigned long long* src_a, unsigned na, const unsigned long long* src_b, unsigned nb,
                                   unsigned long long* dst_host, cudaStream_t stream) {
  k_publish2<<<1, 32, 0, stream>>>(src_a, na, src_b, nb, dst_host);
  return (int)cudaGetLastError();
}

extern "C" int s2m_launch_fp32_probe(int mode, int blocks, int iters, float* sink, cudaStream_t stream) {
  switch (mode) {
    case 0: k_fp32_probe<0><<<blocks, 256, 0, stream>>>(sink, 0.999f, 1e-3f, iters); break;
    case 1: k_fp32_probe<1><<<blocks, 256, 0, stream>>>(sink, 0.999f, 1e-3f, iters); break;
    default: k_fp32_probe<2><<<blocks, 256, 0, stream>>>(sink, 0.999f, 1e-3f, iters); break;
  }
  return (int)cudaGetLastError();
}
/* Corner coordinates of one run, bmin + size * f32(i) per axis (two roundings, as K1 computed them per thread):
 * tab = x[0 .. nx) | y[0 .. ny) | z[0 .. nz).  K1 loads them instead of converting, multiplying and adding. */
__global__ void k_coords(float* __restrict__ tab, unsigned nx, unsigned ny, unsigned nz, float bx, float by, float bz, float sx, float sy, float sz) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nx) tab[i] = __fadd_rn(bx, __fmul_rn(sx, (float)i));
  else if (i < nx + ny) tab[i] = __fadd_rn(by, __fmul_rn(sy, (float)(i - nx)));
  else if (i < nx + ny + nz) tab[i] = __fadd_rn(bz, __fmul_rn(sz, (float)(i - nx - ny)));
}
extern "C" int s2m_launch_coords(float* tab, unsigned nx, unsigned ny, unsigned nz, const float* bmin, const float* size, cudaStream_t stream) {
  const unsigned n = nx + ny + nz;
  k_coords<<<(n + 255u) / 256u, 256, 0, stream>>>(tab, nx, ny, nz, bmin[0], bmin[1], bmin[2], size[0], size[1], size[2]);
  return (int)cudaGetLastError();
}
extern "C" int s2m_launch_publish(const unsigned long long* src, unsigned long long* dst_host, unsigned n, cudaStream_t stream) {
  k_publish<<<1, 32, 0, stream>>>(src, dst_host, n);
  return (int)cudaGetLastError();
}

extern "C" int s2m_launch_k2(const S2mK2Args* a, cudaStream_t stream) {
  dim3 grid((a->res_x + 32 * K2_TX_WORDS - 1) / (32 * K2_TX_WORDS), (a->res_y + K2_TY - 1) / K2_TY,
            (a->nz_chunk + K2_ZT - 1) / K2_ZT);
  if (grid.x == 0 || grid.y == 0 || grid.z == 0) return 0;  /* (the engine never asks for an empty chunk) */
  k2_classify<<<grid, 256, 0, stream>>>(a->slab, a->pitch_x, a->plane_stride, a->res_x, a->res_y, a->nz_chunk, a->tau,
                                        a->cand_mask, a->words_x, a->total, a->done, a->host_total);
  return (int)cudaGetLastError();
}

extern "C" int s2m_launch_k2_bits(const S2mK2Args* a, cudaStream_t stream) {
  dim3 grid((a->words_x + 31u) / 32u, (a->res_y + 7u) / 8u, (a->nz_chunk + K2_ZT - 1) / K2_ZT);
  if (grid.x == 0 || grid.y == 0 || grid.z == 0) return 0;
  k2_classify_bits<<<grid, 256, 0, stream>>>(reinterpret_cast<const uint2*>(a->cls), a->cls_words, a->res_y + 1u, a->res_x, a->res_y, a->nz_chunk,
                                              a->cand_mask, a->words_x, a->total, a->done, a->host_total, a->seg_count);
  return (int)cudaGetLastError();
}

extern "C" uint32_t s2m_segs_x(uint32_t words_x) { return (words_x + 31u) / 32u; }

extern "C" int s2m_launch_seg_count(const uint32_t* cand_mask, unsigned long long n_rows, uint32_t words_x, uint32_t* seg_count, cudaStream_t stream) {
  const uint32_t segs_x = s2m_segs_x(words_x);
  const unsigned long long n_segs = n_rows * segs_x;
  if (!n_segs) return 0;
  k_seg_count<<<(unsigned)((n_segs + 7) / 8), 256, 0, stream>>>(cand_mask, n_rows, words_x, segs_x, seg_count);
  return (int)cudaGetLastError();
}

extern "C" unsigned s2m_k3_tiles(unsigned long long n_words, uint32_t words_x) {
  const unsigned long long n_segs = words_x ? n_words / words_x * s2m_segs_x(words_x) : 0ull;
  return (unsigned)((n_segs + K3_TILE_SEGS - 1) / K3_TILE_SEGS);
}

extern "C" int s2m_launch_k3(const S2mK3Args* a, cudaStream_t stream) {
  const unsigned tiles = s2m_k3_tiles(a->n_words, a->words_x);
  if (!tiles) return 0;
  const uint32_t segs_x = s2m_segs_x(a->words_x);
  k3_compact<<<tiles, K3_THREADS, 0, stream>>>(a->cand_mask, a->seg_count, a->n_words / a->words_x * segs_x, a->words_x, segs_x, a->res_y, a->z_offset,
                                               a->word_prefix, a->cand_key, a->base, a->status, a->ticket);
  return (int)cudaGetLastError();
}

extern "C" unsigned s2m_k4b_tiles(unsigned long long n_own) { return n_own ? (unsigned)((n_own + 255) / 256) : 1u; }

extern "C" int s2m_launch_k4b(const S2mK4bArgs* a, cudaStream_t stream) {
  const unsigned tiles = s2m_k4b_tiles(a->max_vertices);   /* >= 1: an empty chunk still carries the totals forward */
  QuadParams p;
  p.done = a->done; p.host_slot = a->host_slot;
  p.vert_key = a->vert_key; p.vert_nibble = a->vert_nibble; p.tot_prev = a->tot_prev; p.tot_cur = a->tot_cur; p.n_halo = a->n_halo;
  p.cand_mask = a->cand_mask; p.word_prefix = a->word_prefix; p.cand_vrank = a->cand_vrank;
  p.words_x = a->words_x; p.res_y = a->res_y; p.z_first = a->z_first; p.label_add = a->label_add;
  p.index_add = a->index_add; p.quads = a->quads; p.quads32 = a->quads32; p.status = a->status; p.ticket = a->ticket;
  p.n_invalid = a->n_invalid;
  p.invalid_records = a->invalid_records; p.invalid_cursor = a->invalid_cursor; p.invalid_capacity = a->invalid_capacity;
  k4_quads<<<tiles, 256, 0, stream>>>(p);
  return (int)cudaGetLastError();
}
