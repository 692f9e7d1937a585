/* kernels_static.h -- host-callable launchers of the ahead-of-time compiled kernels. */
#ifndef S2M_KERNELS_STATIC_H_
#define S2M_KERNELS_STATIC_H_
#include <cuda_runtime.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct S2mK2Args {
  const float* slab;              /* chunk plane 0 */
  uint32_t pitch_x;
  unsigned long long plane_stride;
  uint32_t res_x, res_y;          /* cells */
  uint32_t nz_chunk;              /* cell slices in this chunk; planes 0..nz_chunk are present */
  float tau;
  uint32_t* cand_mask;            /* first word of the chunk's first slice */
  uint32_t words_x;
  unsigned long long* total;      /* candidate counter (accumulates) */
  const void* cls;                /* s2m_launch_k2_bits: corner-class planes written by K1 (uint2 per 32 corners) */
  uint32_t cls_words;             /* class words per row */
  unsigned* done;                 /* optional: zeroed block-completion counter; the last block to finish ... */
  unsigned long long* host_total; /* ... writes *total here (mapped pinned host memory) */
  uint32_t* seg_count;            /* candidates per SEGMENT (32 consecutive mask words of one cell row, = 1024 cells): entry
                                     (slice * res_y + y) * segs_x + x_word / 32 of the chunk; what K3 scans instead of the mask */
} S2mK2Args;

typedef struct S2mK3Args {
  const uint32_t* cand_mask;
  const uint32_t* seg_count;      /* per-segment candidate counts of the same region (S2mK2Args.seg_count) */
  unsigned long long n_words;
  uint32_t words_x, res_y;
  uint32_t z_offset;              /* true z of slice 0 of the mask */
  uint32_t* word_prefix;          /* same region as cand_mask */
  unsigned long long* cand_key;   /* GLOBAL list: entry `base + local rank` is written */
  unsigned long long base;        /* candidates in earlier chunks */
  unsigned long long* status;     /* >= s2m_k3_tiles(n_words, words_x) zeroed words */
  unsigned* ticket;               /* zeroed */
} S2mK3Args;

typedef struct S2mK4bArgs {
  const unsigned long long* vert_key;
  const unsigned char* vert_nibble;
  unsigned long long max_vertices;   /* upper bound of the vertices of this launch (sizes the grid; the exact range is read on the device) */
  const unsigned long long* tot_prev; /* device: {vertices, quads} emitted up to the end of the previous z-chunk */
  unsigned long long* tot_cur;        /* device: {vertices (written by K4a), quads (written here)} up to the end of this z-chunk */
  const unsigned long long* n_halo;   /* device: vertices of the recomputed slice below the slab (first in the list, no quads of their own) */
  const uint32_t* cand_mask;
  const uint32_t* word_prefix;
  const uint32_t* cand_vrank;
  uint32_t words_x, res_y, z_first, label_add;
  long long index_add;            /* emitted index = local vertex index - n_halo + index_add */
  unsigned long long* quads;
  unsigned* quads32;              /* non-NULL: write 4 x u32 per quad here instead of quads */
  unsigned long long* status;     /* >= s2m_k4b_tiles(max_vertices) zeroed words */
  unsigned* ticket;               /* zeroed */
  unsigned long long* n_invalid;  /* accumulates */
  unsigned long long* invalid_records; /* optional: 6 u64 per invalid quad (key, edge, q0..q3), unordered */
  unsigned long long* invalid_cursor;  /* optional: record counter (accumulates) */
  unsigned long long invalid_capacity;
  unsigned* done;                 /* optional: zeroed block-completion counter; the last block to finish ... */
  unsigned long long* host_slot;  /* ... writes {vertices, quads, n_halo, n_invalid, invalid cursor} here (mapped pinned host memory) */
} S2mK4bArgs;

int s2m_launch_publish(const unsigned long long* src, unsigned long long* dst_host, unsigned n, cudaStream_t stream); /* n <= 32 */
/* dst_host = src_a[0..na) ++ src_b[0..nb), na + nb <= 32 */
int s2m_launch_publish2(const unsigned long long* src_a, unsigned na, const unsigned long long* src_b, unsigned nb, unsigned long long* dst_host,
                        cudaStream_t stream);
/* FP32 throughput probe: 8 FMAs per thread and iteration, 256 threads per block; mode 0 FFMA r,r,r / 1 FFMA r,imm,imm / 2 FFMA2 */
int s2m_launch_fp32_probe(int mode, int blocks, int iters, float* sink, cudaStream_t stream);
int s2m_launch_coords(float* tab, unsigned nx, unsigned ny, unsigned nz, const float* bmin, const float* size, cudaStream_t stream); /* x | y | z corner coordinates for K1 */
int s2m_launch_k2(const S2mK2Args* a, cudaStream_t stream);      /* classify from the f32 slab */
int s2m_launch_k2_bits(const S2mK2Args* a, cudaStream_t stream); /* classify from K1's class bit planes */
uint32_t s2m_segs_x(uint32_t words_x);                                         /* segments per cell row */
/* seg_count of a mask that K2-from-bits did not write (classify from the slab, exact-dense mode) */
int s2m_launch_seg_count(const uint32_t* cand_mask, unsigned long long n_rows, uint32_t words_x, uint32_t* seg_count, cudaStream_t stream);
unsigned s2m_k3_tiles(unsigned long long n_words, uint32_t words_x);
int s2m_launch_k3(const S2mK3Args* a, cudaStream_t stream);
unsigned s2m_k4b_tiles(unsigned long long n_own);
int s2m_launch_k4b(const S2mK4bArgs* a, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif
