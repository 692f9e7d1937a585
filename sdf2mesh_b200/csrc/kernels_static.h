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
} S2mK2Args;

typedef struct S2mK3Args {
  const uint32_t* cand_mask;
  unsigned long long n_words;
  uint32_t words_x, res_y;
  uint32_t z_offset;              /* true z of slice 0 of the mask */
  uint32_t* word_prefix;          /* same region as cand_mask */
  unsigned long long* cand_key;   /* GLOBAL list: entry `base + local rank` is written */
  unsigned long long base;        /* candidates in earlier chunks */
  unsigned long long* status;     /* >= s2m_k3_tiles(n_words) zeroed words */
  unsigned* ticket;               /* zeroed */
} S2mK3Args;

typedef struct S2mK4bArgs {
  const unsigned long long* vert_key;
  const unsigned char* vert_nibble;
  unsigned long long v_begin, v_end; /* vertices to emit quads for (local indices, halo excluded by the caller) */
  unsigned long long quad_base;      /* quads emitted by earlier launches */
  const uint32_t* cand_mask;
  const uint32_t* word_prefix;
  const uint32_t* cand_vrank;
  uint32_t words_x, res_y, z_first, label_add;
  long long index_offset;
  unsigned long long* quads;
  unsigned* quads32;              /* non-NULL: write 4 x u32 per quad here instead of quads */
  unsigned long long* status;     /* >= s2m_k4b_tiles(v_end - v_begin) zeroed words */
  unsigned* ticket;               /* zeroed */
  unsigned long long* n_quads;    /* out: quad_base + quads of this launch */
  unsigned long long* n_invalid;  /* accumulates */
  unsigned long long* invalid_records; /* optional: 6 u64 per invalid quad (key, edge, q0..q3), unordered */
  unsigned long long* invalid_cursor;  /* optional: record counter (accumulates) */
  unsigned long long invalid_capacity;
} S2mK4bArgs;

int s2m_launch_publish(const unsigned long long* src, unsigned long long* dst_host, unsigned n, cudaStream_t stream); /* n <= 32 */
int s2m_launch_k2(const S2mK2Args* a, cudaStream_t stream);      /* classify from the f32 slab */
int s2m_launch_k2_bits(const S2mK2Args* a, cudaStream_t stream); /* classify from K1's class bit planes */
unsigned s2m_k3_tiles(unsigned long long n_words);
int s2m_launch_k3(const S2mK3Args* a, cudaStream_t stream);
unsigned s2m_k4b_tiles(unsigned long long n_own);
int s2m_launch_k4b(const S2mK4bArgs* a, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif
