/* kernels_jit.cuh -- the kernels that call the user's SDF; compiled per module with NVRTC for
 * sm_100a (--fmad=false) after: s2m_math.h, s2m_vec.h, s2m_sdf3d_lib.h, s2m_scan.cuh, and the CUDA C++
 * the front-end emitted for the user's `fn sdf3d(p: vec3f) -> f32` (namespace s2m_user).
 *
 * K1  s2m_k1_slab      replaces cell_new's 8 evaluations per cell (/root/reference/src/bin/sdf2mesh/
 *                      dualcontour.wgsl:29-43) with ONE evaluation per grid corner, written with
 *                      coalesced float4 stores into the HBM-resident slab.
 * K4a s2m_k4_vertices  replaces, for the candidate cells only, the whole shader entry point
 *                      (dualcontour.wgsl:161-180): cell_bounds :22-27, cell_new :29-43 with the
 *                      reference's exact per-cell corner arithmetic (min, min+size), the 12-edge
 *                      crossing average :86-131, sdf3d_normal (sdf3d_normal.wgsl:4-10) + normalize,
 *                      the sign nibble :57-69 -- and the host pixel scan + VertexList::insert
 *                      (main.rs:327-344, mesh.rs:237) by a stable single-pass compaction.
 *     s2m_k_eval       evaluates the SDF at caller-supplied points (diagnostics / parity tests).
 *     s2m_k_cost_probe per-z-plane evaluation cost from a lattice of scalar evaluations (S2M_COST_PROBE=lattice; by default
 *                      s2m_cost_probe times K1 itself on one plane per z-band, engine.cpp).
 */

/* The module is compiled as three independent NVRTC programs, concurrently on three host threads
 * (engine.cpp build_module): S2M_JIT_PART = 1: K1, 2: K4a, 3: the diagnostic kernels; undefined: all
 * of them in one translation unit (tools/, offline nvcc).  Each part carries the copies of the
 * user's SDF that it needs and nothing else, which is what the compile time is made of. */
#if !defined(S2M_JIT_PART)
#define S2M_JIT_K1 1
#define S2M_JIT_K4 1
#define S2M_JIT_MISC 1
#elif S2M_JIT_PART == 1
#define S2M_JIT_K1 1
#elif S2M_JIT_PART == 2
#define S2M_JIT_K4 1
#elif S2M_JIT_PART == 3
#define S2M_JIT_MISC 1
#endif

#ifndef S2M_K1_UNROLL
#define S2M_K1_UNROLL 4
#endif

struct S2mGrid {
  float bmin[3];
  float size[3];      /* (bmax - bmin) / f32(res - 1), computed on the host in f32 (IEEE division) */
  float eps;
  unsigned res[3];    /* cells per axis; corners per axis = res + 1 */
  unsigned pitch_x;   /* floats per slab row (multiple of 32) */
  unsigned rows;      /* res[1] + 1 */
  unsigned long long plane_stride; /* pitch_x * rows */
};

__device__ __forceinline__ float s2m_sdf(float x, float y, float z) {
  return s2m_user::sdf3d(s2m::mk3(x, y, z));
}
__device__ __noinline__ float s2m_sdf_call(float x, float y, float z) { return s2m_sdf(x, y, z); }

/* ------------------------------------------------------------------------------------------ K1 */
#if defined(S2M_JIT_K1)
/* A thread evaluates 4 consecutive x corners and stores one float4.  The block shape is chosen by
 * the host: blockDim = (bx, by) with bx*by = 256.  A warp is 32 consecutive threads, x fastest, so
 * its footprint is (4*min(bx,32)) x (32/min(bx,32)) corners: (32,8) = 128x1 rows (512 B contiguous
 * per warp), (8,32) = 32x4 tiles (one full 128 B line per row) -- more compact, less divergence
 * in SDFs whose cost varies in space.  grid = (ceil(pitch_x/(4*bx)), ceil(rows/by), planes). */
#ifndef S2M_K1_ROWS
#define S2M_K1_ROWS 2   /* grid rows a thread evaluates (1 or 2): 2 amortises the index / class / store overhead */
#endif

/* `redo`: with S2M_K1_PACKED the corners (0,1) and (2,3) are evaluated as f32x2 pairs
 * (s2m_user_p::sdf3d2, s2m_pvec.h); bit (redo_shift + k/2) is set when the lanes of pair k disagreed on a
 * comparison or conversion, i.e. v[k+1] is not valid and the caller evaluates that corner alone. */
__device__ __forceinline__ void s2m_k1_eval4(bool on, const float cx[4], float cy, float cz, float v[4], unsigned& redo, unsigned redo_shift) {
  /* Without S2M_K1_GUARD every thread that is still here evaluates its 4 corners, also those in the row padding past
   * the last corner (x4 > res; their coordinates continue the grid) and, with two rows per thread, a row past the last
   * one -- they sit in warps that run for their in-grid lanes anyway, nobody reads what they produce (K2 masks the cells
   * beyond the grid), and the guard with its zero fill is 6 instructions per warp (torus 2048^3 K1 7.5 -> 6.95 ms).
   * S2M_K1_GUARD=1 is an experiment knob: with one plane per thread the mandelbulb's K1 was 1 ms slower without the
   * guard (the compiler paired registers differently), with its plane loop it is 0.4 ms faster
   * (profiles/r02_k1_ab.jsonl). */
#if defined(S2M_K1_GUARD)
  v[0] = 0.0f; v[1] = 0.0f; v[2] = 0.0f; v[3] = 0.0f;
  if (on)
#else
  (void)on;
#endif
  {
#if S2M_K1_UNROLL == 1
#pragma unroll 1
#else
#pragma unroll
#endif
#if defined(S2M_K1_PACKED)
    for (int k = 0; k < 4; k += 2) {
      bool dv;
      const s2m::pf r = s2m_user_p::sdf3d2(s2m::pmk3(s2m::pf(cx[k], cx[k + 1]), s2m::pf(cy), s2m::pf(cz)), &dv);
      v[k] = r.lo; v[k + 1] = r.hi;
      if (dv) redo |= 1u << (redo_shift + (unsigned)(k >> 1));
    }
#else
    for (int k = 0; k < 4; ++k) v[k] = s2m_sdf(cx[k], cy, cz);
    (void)redo; (void)redo_shift;
#endif
  }
}
/* corner classes of 4 values: low nibble P (value > tau), high nibble N (value < -tau).
 * Sign bits of differences instead of compares: v > tau <=> tau - v < 0 and v < -tau <=> v + tau < 0 -- a difference of two
 * floats that is not zero keeps its sign when rounded, an exact zero is +0, and a NaN operand gives the canonical NaN
 * 0x7fffffff (sign clear: neither class, as with the compares).  One FADD (FMA pipe) and one funnel shift per bit instead
 * of FSETP + SEL + LOP3 on the half-rate ALU pipe, which is what the tiny SDFs' K1 runs out of (torus: 29 instructions
 * per corner, a third of them this epilogue). */
__device__ __forceinline__ unsigned s2m_k1_class_byte(const float v[4], float tau) {
  unsigned acc = 0u;
#pragma unroll
  for (int k = 3; k >= 0; --k) acc = __funnelshift_l(__float_as_uint(v[k] + tau), acc, 1);   /* N3 .. N0 -> bits 7 .. 4 */
#pragma unroll
  for (int k = 3; k >= 0; --k) acc = __funnelshift_l(__float_as_uint(tau - v[k]), acc, 1);   /* P3 .. P0 -> bits 3 .. 0 */
  return acc;
}

#ifndef S2M_K1_ZPT
#define S2M_K1_ZPT 1   /* > 1: a thread marches through `zpt` planes (a launch parameter <= S2M_K1_ZPT; grid.z = ceil(planes / zpt)) */
#endif
#ifndef S2M_K1_MINBLOCKS
#define S2M_K1_MINBLOCKS 1   /* resident 256-thread blocks per SM the register allocator must allow (8 = at most 32 registers) */
#endif
/* Consecutive z-chunks share one corner plane: the top plane of the previous chunk (carry_*: that plane in the
 * previous chunk's buffers) is this launch's plane 0.  Its blocks copy it instead of evaluating it again -- same
 * values, a plane's worth of SDF evaluations saved per chunk boundary (3 % of K1 for a 146-slice slab in 6 chunks).
 * Kept out of the plane loop: inside it the copy cost the tiny kernels 11 registers (torus 58 -> 69).  Takes the pitch
 * by value: a reference to the grid struct made every thread spill the kernel's parameters to local memory at entry
 * (7 STL per thread, 1.7 % of the mandelbulb K1's instructions; ncu source view). */
__device__ __noinline__ void s2m_k1_carry_plane(unsigned pitch_x, float* __restrict__ slab, uint2* __restrict__ cls,
                                                const float* __restrict__ carry_slab, const uint2* __restrict__ carry_cls,
                                                unsigned x4, unsigned y, bool active, bool active_b) {
  if (slab != nullptr && carry_slab != nullptr) {
    if (active) *reinterpret_cast<float4*>(slab + (unsigned long long)y * pitch_x + x4) = __ldg(reinterpret_cast<const float4*>(carry_slab + (unsigned long long)y * pitch_x + x4));
    if (active_b) *reinterpret_cast<float4*>(slab + (y + 1ull) * pitch_x + x4) = __ldg(reinterpret_cast<const float4*>(carry_slab + (y + 1ull) * pitch_x + x4));
  }
  if (cls != nullptr && carry_cls != nullptr) {   /* every thread copies the class byte(s) it stores in the epilogue */
    const unsigned long long at = (unsigned long long)y * (pitch_x >> 2) + (x4 >> 2);
    unsigned char* dst = reinterpret_cast<unsigned char*>(cls);
    const unsigned char* src = reinterpret_cast<const unsigned char*>(carry_cls);
    if (active) dst[at] = __ldg(src + at);
    if (active_b) dst[at + (pitch_x >> 2)] = __ldg(src + at + (pitch_x >> 2));
  }
}

extern "C" __global__ void __launch_bounds__(256, S2M_K1_MINBLOCKS)
s2m_k1_slab(S2mGrid g, float* __restrict__ slab, unsigned first_plane, unsigned n_planes,
            float tau, uint2* __restrict__ cls, unsigned cls_words,
            const float* __restrict__ carry_slab, const uint2* __restrict__ carry_cls,
            const float* __restrict__ coord_x, const float* __restrict__ coord_y, const float* __restrict__ coord_z, unsigned opt, unsigned zpt) {
  /* opt: what the pointer arguments say, as bits (one 32-bit test each instead of 64-bit pointer compares per thread):
   * 1 = carry_slab or carry_cls is set, 2 = slab is set, 4 = cls is set.  coord_z points at first_plane's entry. */
  const unsigned x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4u;
  const unsigned y = (blockIdx.y * blockDim.y + threadIdx.y) * (unsigned)S2M_K1_ROWS;
  /* threads past the padded row or the last row have nothing to do (no warp-wide operation below needs them) */
  if (x4 >= g.pitch_x || y >= g.rows) return;
  const bool active = true;
  /* S2M_K1_UNROLL=1 keeps ONE inlined copy of the SDF per row in the kernel (4x less code); 4 =
   * unrolled (lets independent evaluations overlap; faster for the mandelbulb, measured). */
  float cx[4];
#if defined(S2M_K1_COORDS)
  /* coordinates from the run's table (k_coords: the same two roundings): one 16-byte load for the 4 x, one load each for
   * y and z, instead of a conversion, a multiplication and an addition per coordinate -- and nothing cheap for the
   * register allocator to compute again in front of every inlined SDF (ncu source view: 45 of the mandelbulb K1's
   * ~690 instructions per warp were coordinates) */
  {
    const float4 t = __ldg(reinterpret_cast<const float4*>(coord_x + x4));
    cx[0] = t.x; cx[1] = t.y; cx[2] = t.z; cx[3] = t.w;
  }
  const float cy_a = __ldg(coord_y + y);
#else
  (void)coord_x; (void)coord_y; (void)coord_z;
#pragma unroll
  for (int k = 0; k < 4; ++k) cx[k] = g.bmin[0] + g.size[0] * (float)(x4 + (unsigned)k);
  const float cy_a = g.bmin[1] + g.size[1] * (float)y;
#endif
#if S2M_K1_ROWS == 2
  const bool active_b = active && y + 1u < g.rows;
#if defined(S2M_K1_COORDS)
  const float cy_b = __ldg(coord_y + y + 1u);
#else
  const float cy_b = g.bmin[1] + g.size[1] * (float)(y + 1u);
#endif
#endif
  /* For a tiny SDF a thread marches through S2M_K1_ZPT consecutive planes: its x and y coordinates, its indices and its
   * activity are computed once (they are a third of the instructions of a torus evaluation), only z changes.  Not
   * unrolled: one inlined copy of the SDF per row either way.  The plane range is the same for the whole block.
   * Measured on B200 (profiles/r02_k1_ab.jsonl): torus 2048^3 K1 11.4 -> 8.7 ms at 16 planes.  The mandelbulb lost
   * 1-2 % while its per-thread overhead was hidden among 48-register spills and shuffles; after the round-2 diet it
   * gains 8 % (K1 35.5 -> 32.6 ms at 16 planes).  The large primitive compositions keep one plane per thread. */
  const bool carry = blockIdx.z == 0u && (opt & 1u) != 0u;   /* uniform over the block */
#if S2M_K1_ROWS == 2
  if (carry) s2m_k1_carry_plane(g.pitch_x, slab, cls, carry_slab, carry_cls, x4, y, active, active_b);
#else
  if (carry) s2m_k1_carry_plane(g.pitch_x, slab, cls, carry_slab, carry_cls, x4, y, active, false);
#endif
#if S2M_K1_ZPT > 1
  const unsigned pz_end = min(n_planes, (blockIdx.z + 1u) * zpt);
#pragma unroll 1
  for (unsigned pz = blockIdx.z * zpt + (carry ? 1u : 0u); pz < pz_end; ++pz) {
#else
  (void)zpt;
  if (!carry) {  /* one plane per thread: no loop (a loop of one iteration still costs the larger kernels registers) */
  const unsigned pz = blockIdx.z;
#endif
  /* the thread's first corner as a float index into the chunk's slab; a launch has fewer than 2^32 rows.  The class
   * byte of 4 corners sits at a quarter of it (pitch_x and x4 are multiples of 4). */
  const unsigned long long idx = (unsigned long long)(pz * g.rows + y) * g.pitch_x + x4;
#if defined(S2M_K1_COORDS)
  const float cz = __ldg(coord_z + pz);
#else
  const float cz = g.bmin[2] + g.size[2] * (float)(first_plane + pz);
#endif
  float va[4];
  unsigned redo = 0;
  s2m_k1_eval4(x4 <= g.res[0], cx, cy_a, cz, va, redo, 0u);
#if S2M_K1_ROWS == 2
  float vb[4];
  s2m_k1_eval4(active_b && x4 <= g.res[0], cx, cy_b, cz, vb, redo, 2u);
#endif
#if defined(S2M_K1_PACKED)
  /* Rare (0.3 % of the mandelbulb's pairs at 2048^3): the two lanes of a pair took different paths.
   * Lane lo is always right; the hi corner is evaluated again on its own, all such corners of the
   * thread in one loop so that a warp pays max-over-lanes evaluations, not one per call site. */
  while (redo) {
    const unsigned b = (unsigned)__ffs((int)redo) - 1u;
    redo &= redo - 1u;
#if S2M_K1_ROWS == 2
    const float r = s2m_sdf_call((b & 1u) ? cx[3] : cx[1], (b & 2u) ? cy_b : cy_a, cz);
    if (b == 0u) va[1] = r; else if (b == 1u) va[3] = r; else if (b == 2u) vb[1] = r; else vb[3] = r;
#else
    const float r = s2m_sdf_call((b & 1u) ? cx[3] : cx[1], cy_a, cz);
    if (b == 0u) va[1] = r; else va[3] = r;
#endif
  }
#endif
  /* slab == nullptr: the slab-free form for cheap SDFs (S2M_MESH_NO_SLAB) -- only the corner classes below are
   * written (0.25 B per corner instead of 4.25) and K4a evaluates all 8 corners of every candidate cell itself. */
  if (opt & 2u) {
    if (active) *reinterpret_cast<float4*>(slab + idx) = make_float4(va[0], va[1], va[2], va[3]);
#if S2M_K1_ROWS == 2
    if (active_b) *reinterpret_cast<float4*>(slab + idx + g.pitch_x) = make_float4(vb[0], vb[1], vb[2], vb[3]);
#endif
  }
  /* Corner classes for K2: P = value > +tau, N = value < -tau (NaN and the |v| <= tau band are
   * neither).  One byte per thread and row: low nibble = P of its 4 corners, high nibble = N; byte l of
   * cls[plane][row][x/32] holds corners 4l .. 4l+3, i.e. the bytes of a row are the threads of a row in order:
   * 0.25 B per corner instead of K2 re-reading 4 B.  Every thread stores its own byte -- a warp's 32 bytes are one
   * 32-byte sector -- which is 15 instructions per warp less than assembling words with two shuffles and storing
   * them from every fourth lane (5 % of the instructions of a mandelbulb warp outside the fractal). */
  if (opt & 4u) {
    unsigned char* cb = reinterpret_cast<unsigned char*>(cls) + (idx >> 2);
    if (active) *cb = (unsigned char)s2m_k1_class_byte(va, tau);
#if S2M_K1_ROWS == 2
    if (active_b) cb[g.pitch_x >> 2] = (unsigned char)s2m_k1_class_byte(vb, tau);
#endif
  }
  }  /* planes */
}

#endif  /* S2M_JIT_K1 */

/* ------------------------------------------------------------------------------------------ K4a */
#if defined(S2M_JIT_K4)
__device__ __forceinline__ float s2m_cell_adapt(float v0, float v1) { return (0.0f - v0) / (v1 - v0); }

struct S2mVertexOut {
  float* pos;                 /* 3 per vertex */
  float* nrm;                 /* 3 per vertex */
  unsigned long long* key;    /* x | y<<16 | label<<32  (mesh.rs:224-226) */
  unsigned char* nibble;      /* bit0 s100, bit1 s010, bit2 s001, bit3 s000 */
  unsigned* cand_vrank;       /* per candidate: vertex index or 0xffffffff */
  unsigned long long* status; /* look-back tile status, zeroed by the host */
  unsigned* ticket;           /* tile ticket counter, zeroed by the host */
  unsigned long long* n_vertices;  /* out: vertices emitted up to the end of this launch (vert_base + this launch's) */
  unsigned long long* n_halo;      /* out: vertices whose true z < halo_below */
};

/* the part of K1's slab that is still resident when K4a runs (n_planes = 0: none) */
struct S2mSlabView {
  const float* slab;
  unsigned first_plane;
  unsigned n_planes;
};

#define S2M_K4_THREADS 128

/* Distribute `count` work items per lane over the lanes of a warp: returns the warp total and
 * writes the inclusive prefix to inc[lane] (shared memory, 32 entries per warp). */
__device__ __forceinline__ unsigned s2m_warp_items(unsigned count, unsigned* inc) {
  const unsigned lane = threadIdx.x & 31u;
  unsigned v = count;
  for (int o = 1; o < 32; o <<= 1) {
    unsigned n = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= (unsigned)o) v += n;
  }
  inc[lane] = v;
  __syncwarp();
  return __shfl_sync(0xffffffffu, v, 31);
}
/* owner lane of work item `item`: first lane whose inclusive prefix exceeds it */
__device__ __forceinline__ unsigned s2m_item_owner(const unsigned* inc, unsigned item) {
  unsigned lo = 0;
#pragma unroll
  for (int step = 16; step > 0; step >>= 1)
    if (inc[lo + step - 1] <= item) lo += step;
  return lo;
}

/* One thread per candidate cell; SDF evaluations are pooled per warp and dealt out evenly.
 *
 * The reference evaluates the 8 corners of a cell at (min | min+size) per axis
 * (dualcontour.wgsl:22-43).  Corner 000 is always at the slab's own coordinates, and a corner with
 * a `max` coordinate is too whenever fl(min + size) == the next corner's min bit for bit (true for
 * 17-95 % of indices).  Those values are read back from K1's slab (same function, same
 * arguments, same bits); only the rest -- 2.5 of 8 on average at 2048^3 -- are evaluated, and those
 * evaluations plus the 4 normal taps of cells that get a vertex are spread over all 32 lanes.
 *
 * cand_key = x | y<<16 | z_true<<32 (this launch's slice of the list; cand_vrank likewise).
 * *vert_base_ptr = vertices emitted by earlier chunks.  label_add = 1 in faithful mode.
 * mode: bit 0 = compute normals, bit 1 = consistent corners. */
extern "C" __global__ void __launch_bounds__(S2M_K4_THREADS)
s2m_k4_vertices(S2mGrid g, const unsigned long long* __restrict__ cand_key, unsigned long long n_cand,
                const unsigned long long* __restrict__ vert_base_ptr, unsigned label_add, unsigned halo_below, unsigned mode, S2mSlabView sv, S2mVertexOut out) {
  /* vertices emitted by earlier z-chunks: read from device memory (the previous launch's out.n_vertices), so that
   * the host does not have to learn it between two chunks */
  const unsigned long long vert_base = *vert_base_ptr;
  const unsigned want_normals = mode & 1u;
  const bool consistent = (mode & 2u) != 0u;  /* S2M_MESH_CONSISTENT_CORNERS: a cell's max corner IS the next cell's min corner */
  __shared__ unsigned s_scan[33];
  __shared__ unsigned s_tile;
  __shared__ unsigned long long s_base;
  __shared__ unsigned s_inc[S2M_K4_THREADS];        /* per warp: inclusive item prefix */
  __shared__ unsigned s_mask[S2M_K4_THREADS];       /* per lane: which corners / taps need an evaluation */
  __shared__ float s_co[S2M_K4_THREADS][6];         /* per lane: cmin.xyz, cmax.xyz  (later: pos.xyz) */
  __shared__ float s_val[S2M_K4_THREADS][8];        /* per lane: evaluated corner values / tap values */
  if (threadIdx.x == 0) s_tile = atomicAdd(out.ticket, 1u);
  __syncthreads();
  const unsigned tile = s_tile;
  const unsigned long long c = (unsigned long long)tile * blockDim.x + threadIdx.x;
  const unsigned lane = threadIdx.x & 31u;
  const unsigned wbase = threadIdx.x & ~31u;          /* first thread of this warp */
  unsigned* inc = s_inc + wbase;

  bool has_vertex = false;
  float pos[3] = {0.f, 0.f, 0.f}, nrm[3] = {0.f, 0.f, 0.f};
  float cmin[3] = {0.f, 0.f, 0.f}, cmax[3] = {0.f, 0.f, 0.f};
  float d[8];
  unsigned nib = 0, need = 0;
  unsigned cx = 0, cy = 0, cz = 0;
  const bool live = c < n_cand;
  if (live) {
    const unsigned long long key = cand_key[c];
    cx = (unsigned)(key & 0xffffu); cy = (unsigned)((key >> 16) & 0xffffu); cz = (unsigned)(key >> 32);
    /* cell_bounds, dualcontour.wgsl:22-27 */
    const float nx = g.bmin[0] + g.size[0] * (float)(cx + 1u), ny = g.bmin[1] + g.size[1] * (float)(cy + 1u), nz = g.bmin[2] + g.size[2] * (float)(cz + 1u);
    cmin[0] = g.bmin[0] + g.size[0] * (float)cx; cmax[0] = consistent ? nx : cmin[0] + g.size[0];
    cmin[1] = g.bmin[1] + g.size[1] * (float)cy; cmax[1] = consistent ? ny : cmin[1] + g.size[1];
    cmin[2] = g.bmin[2] + g.size[2] * (float)cz; cmax[2] = consistent ? nz : cmin[2] + g.size[2];
    /* does the reference's `max` coordinate coincide with the slab coordinate of the next corner? */
    const bool mx = cmax[0] == nx, my = cmax[1] == ny, mz = cmax[2] == nz;
    const bool p0 = cz >= sv.first_plane && cz - sv.first_plane < sv.n_planes;          /* plane cz resident */
    const bool p1 = cz + 1u >= sv.first_plane && cz + 1u - sv.first_plane < sv.n_planes; /* plane cz+1 resident */
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const bool ok = ((k & 1) ? mx : true) && ((k & 2) ? my : true) && ((k & 4) ? (mz && p1) : p0);
      if (ok) {
        const unsigned long long at = (unsigned long long)(cz + ((k >> 2) & 1) - sv.first_plane) * g.plane_stride +
                                      (unsigned long long)(cy + ((k >> 1) & 1)) * g.pitch_x + (cx + (k & 1));
        d[k] = __ldg(sv.slab + at);
      } else {
        d[k] = 0.0f;
        need |= 1u << k;
      }
    }
  }
  /* ---- pooled corner evaluations: cell_new :29-43, the reference's own 8 corner positions */
  s_mask[threadIdx.x] = need;
#pragma unroll
  for (int a = 0; a < 3; ++a) { s_co[threadIdx.x][a] = cmin[a]; s_co[threadIdx.x][3 + a] = cmax[a]; }
  {
    const unsigned total = s2m_warp_items((unsigned)__popc(need), inc);
    for (unsigned item = lane; item < total; item += 32u) {
      const unsigned o = s2m_item_owner(inc, item);
      const unsigned j = item - (o ? inc[o - 1] : 0u);
      const unsigned k = __fns(s_mask[wbase + o], 0u, (int)j + 1);
      const float* co = s_co[wbase + o];
      s_val[wbase + o][k] = s2m_sdf_call((k & 1u) ? co[3] : co[0], (k & 2u) ? co[4] : co[1], (k & 4u) ? co[5] : co[2]);
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (need & (1u << k)) d[k] = s_val[threadIdx.x][k];
    __syncwarp();
  }
  if (live) {
    /* cell_fetch_interpolated_pos :86-131.  Edge e joins corners ea[e] -> eb[e]; the crossing
     * parameter goes into axis ax[e]; the other two coordinates are the corner bits of ea[e]. */
    const int ea[12] = {0, 2, 1, 3, 0, 4, 1, 5, 0, 4, 2, 6};
    const int eb[12] = {4, 6, 5, 7, 2, 6, 3, 7, 1, 5, 3, 7};
    const int ax[12] = {2, 2, 2, 2, 1, 1, 1, 1, 0, 0, 0, 0};
    float avg[3] = {0.0f, 0.0f, 0.0f};
    float count = 0.0f;
#pragma unroll
    for (int e = 0; e < 12; ++e) {
      const float v0 = d[ea[e]], v1 = d[eb[e]];
      if ((v0 > 0.0f) != (v1 > 0.0f)) {
        float ch[3];
        ch[0] = (ea[e] & 1) ? 1.0f : 0.0f; ch[1] = (ea[e] & 2) ? 1.0f : 0.0f; ch[2] = (ea[e] & 4) ? 1.0f : 0.0f;
        ch[ax[e]] = s2m_cell_adapt(v0, v1);
        if (ch[0] > 0.0f || ch[1] > 0.0f || ch[2] > 0.0f) {
          avg[0] += ch[0]; avg[1] += ch[1]; avg[2] += ch[2];
          count += 1.0f;
        }
      }
    }
    if (!(count <= 1.0f)) {
      has_vertex = true;
      pos[0] = cmin[0] + (cmax[0] - cmin[0]) * avg[0] / count;
      pos[1] = cmin[1] + (cmax[1] - cmin[1]) * avg[1] / count;
      pos[2] = cmin[2] + (cmax[2] - cmin[2]) * avg[2] / count;
      nib = (d[1] > 0.0f ? 1u : 0u) | (d[2] > 0.0f ? 2u : 0u) | (d[4] > 0.0f ? 4u : 0u) | (d[0] > 0.0f ? 8u : 0u);
    }
  }
  /* stable compaction, first half: the tile's vertex count is known now -- announce it before the normal taps (60 % of
   * this kernel's evaluations), look back for the prefix after them: the tiles' look-back then finds its predecessors
   * published instead of spinning on them (17 % of K4a's instructions were that spin; ncu source view, round 2) */
  unsigned total = 0;
  const unsigned local = s2m_block_exclusive_scan(has_vertex ? 1u : 0u, s_scan, &total);
  if (threadIdx.x == 0) s2m_publish_aggregate(out.status, tile, (unsigned long long)total, vert_base);
  /* ---- pooled normal taps: sdf3d_normal.wgsl:4-10
   *      v1*f(p+v1*eps) + v2*f(p+v2*eps) + v3*f(p+v3*eps) + v4*f(p+v4*eps), then normalize (:171) */
  if (want_normals) {
    const float v[4][3] = {{1.0f, -1.0f, -1.0f}, {-1.0f, -1.0f, 1.0f}, {-1.0f, 1.0f, -1.0f}, {1.0f, 1.0f, 1.0f}};
    s_co[threadIdx.x][0] = pos[0]; s_co[threadIdx.x][1] = pos[1]; s_co[threadIdx.x][2] = pos[2];
    const unsigned total = s2m_warp_items(has_vertex ? 4u : 0u, inc);
    for (unsigned item = lane; item < total; item += 32u) {
      const unsigned o = s2m_item_owner(inc, item);
      const unsigned k = item - (o ? inc[o - 1] : 0u);
      const float* p = s_co[wbase + o];
      const float sx = (k == 0u || k == 3u) ? 1.0f : -1.0f, sy = (k >= 2u) ? 1.0f : -1.0f, sz = (k == 1u || k == 3u) ? 1.0f : -1.0f;
      s_val[wbase + o][k] = s2m_sdf_call(p[0] + sx * g.eps, p[1] + sy * g.eps, p[2] + sz * g.eps);
    }
    __syncwarp();
    if (has_vertex) {
      float n[3];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float f = s_val[threadIdx.x][k];
        if (k == 0) { n[0] = v[k][0] * f; n[1] = v[k][1] * f; n[2] = v[k][2] * f; }
        else { n[0] = n[0] + v[k][0] * f; n[1] = n[1] + v[k][1] * f; n[2] = n[2] + v[k][2] * f; }
      }
      const float len = sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
      nrm[0] = n[0] / len; nrm[1] = n[1] / len; nrm[2] = n[2] / len;
    }
  }
  /* stable compaction, second half: decoupled look-back for the tile's base */
  if (threadIdx.x < 32) {
    unsigned long long b = s2m_lookback_published(out.status, tile, (unsigned long long)total, vert_base);
    if (threadIdx.x == 0) s_base = b;
  }
  __syncthreads();
  const unsigned long long vi = s_base + local;
  if (live) out.cand_vrank[c] = has_vertex ? (unsigned)vi : 0xffffffffu;
  if (has_vertex) {
    out.pos[3 * vi + 0] = pos[0]; out.pos[3 * vi + 1] = pos[1]; out.pos[3 * vi + 2] = pos[2];
    out.nrm[3 * vi + 0] = nrm[0]; out.nrm[3 * vi + 1] = nrm[1]; out.nrm[3 * vi + 2] = nrm[2];
    out.key[vi] = (unsigned long long)cx | ((unsigned long long)cy << 16) | ((unsigned long long)(cz + label_add) << 32);
    out.nibble[vi] = (unsigned char)nib;
    if (cz < halo_below) atomicAdd(out.n_halo, 1ull);
  }
  /* the last tile publishes the total */
  if (threadIdx.x == 0 && (unsigned long long)(tile + 1) * blockDim.x >= n_cand) *out.n_vertices = s_base + total;
}

#endif  /* S2M_JIT_K4 */

/* ------------------------------------------------------------------------------------------ misc */
#if defined(S2M_JIT_MISC)
extern "C" __global__ void s2m_k_eval(const float* __restrict__ pts, float* __restrict__ out, unsigned long long n) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = s2m_sdf(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
}

#if defined(S2M_K1_PACKED)
/* the packed form on caller-supplied pairs, raw: lane lo = a[i], lane hi = b[i]; dv[i] = the lanes
 * disagreed (out_b[i] is then not valid).  Parity tests compare this with s2m_k_eval. */
extern "C" __global__ void s2m_k_eval2(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out_a,
                                       float* __restrict__ out_b, unsigned char* __restrict__ dv, unsigned long long n) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  bool d;
  const s2m::pf r = s2m_user_p::sdf3d2(s2m::pmk3(s2m::pf(a[3 * i], b[3 * i]), s2m::pf(a[3 * i + 1], b[3 * i + 1]), s2m::pf(a[3 * i + 2], b[3 * i + 2])), &d);
  out_a[i] = r.lo; out_b[i] = r.hi; dv[i] = d ? 1 : 0;
}
#endif

/* Coarse cost probe: block b evaluates a (probe x probe) lattice of plane z_b of a `planes`-plane
 * coarse grid; every warp adds the SM cycles it spent (issue time including divergence), so the
 * per-plane sums are proportional to K1's work per z-slice. */
extern "C" __global__ void __launch_bounds__(256)
s2m_k_cost_probe(S2mGrid g, unsigned probe, unsigned planes, unsigned long long* __restrict__ cycles, float* __restrict__ sink) {
  const unsigned pz = blockIdx.x;
  const float fz = g.bmin[2] + (g.size[2] * (float)g.res[2]) * ((float)pz + 0.5f) / (float)planes;
  float acc = 0.0f;
  /* a warp covers a compact 8x4 patch per step, like K1's warps */
  const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  const unsigned tiles_x = (probe + 7u) / 8u, tiles_y = (probe + 3u) / 4u;
  const long long t0 = clock64();
  for (unsigned t = warp; t < tiles_x * tiles_y; t += blockDim.x >> 5) {
    const unsigned ix = (t % tiles_x) * 8u + (lane & 7u), iy = (t / tiles_x) * 4u + (lane >> 3);
    if (ix < probe && iy < probe) {
      const float fx = g.bmin[0] + (g.size[0] * (float)g.res[0]) * ((float)ix + 0.5f) / (float)probe;
      const float fy = g.bmin[1] + (g.size[1] * (float)g.res[1]) * ((float)iy + 0.5f) / (float)probe;
      acc += s2m_sdf(fx, fy, fz);
    }
  }
  const long long t1 = clock64();
  if (lane == 0) atomicAdd(cycles + pz, (unsigned long long)(t1 - t0));
  if (acc == 123.456f) sink[0] = acc;
}
#endif  /* S2M_JIT_MISC */
