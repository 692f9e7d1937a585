// Microbenchmark: FP32 throughput of FFMA vs FFMA2 on sm_100a, alone and mixed with integer work.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o ffma2_rate ffma2_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 8192
#define CHAINS 8

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float b, float c, int ib) {
  float a[CHAINS];
  float2 p[CHAINS / 2];
  unsigned u[CHAINS];
  for (int i = 0; i < CHAINS; ++i) { a[i] = threadIdx.x * 1e-3f + i; u[i] = threadIdx.x + i; }
  for (int i = 0; i < CHAINS / 2; ++i) p[i] = make_float2(a[2 * i], a[2 * i + 1]);
  const float2 b2 = make_float2(b, b * 1.0001f), c2 = make_float2(c, c * 0.999f), bb = make_float2(b, b);
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
    if (MODE == 0) {           // FFMA reg,reg,reg
#pragma unroll
      for (int i = 0; i < CHAINS; ++i) a[i] = fmaf(a[i], b, c);
    } else if (MODE == 1) {    // FFMA with immediates
#pragma unroll
      for (int i = 0; i < CHAINS; ++i) a[i] = fmaf(a[i], 0.999f, 1e-3f);
    } else if (MODE == 2) {    // FFMA2 pair,pair,pair  (CHAINS/2 instructions = CHAINS fmas)
#pragma unroll
      for (int i = 0; i < CHAINS / 2; ++i) p[i] = __ffma2_rn(p[i], b2, c2);
    } else if (MODE == 3) {    // FFMA2 pair, broadcast scalar, immediate
#pragma unroll
      for (int i = 0; i < CHAINS / 2; ++i) p[i] = __ffma2_rn(p[i], bb, make_float2(1e-3f, 1e-3f));
    } else if (MODE == 4) {    // mix: 8 FFMA + 8 integer ops
#pragma unroll
      for (int i = 0; i < CHAINS; ++i) { a[i] = fmaf(a[i], b, c); u[i] = (u[i] ^ (unsigned)ib) + (u[i] >> 3); }
    } else if (MODE == 5) {    // mix: 4 FFMA2 (= 8 fmas) + 8 integer ops
#pragma unroll
      for (int i = 0; i < CHAINS / 2; ++i) p[i] = __ffma2_rn(p[i], b2, c2);
#pragma unroll
      for (int i = 0; i < CHAINS; ++i) u[i] = (u[i] ^ (unsigned)ib) + (u[i] >> 3);
    } else if (MODE == 6) {    // 2x the chains in FFMA2: 8 FFMA2 = 16 fmas
#pragma unroll
      for (int i = 0; i < CHAINS / 2; ++i) { p[i] = __ffma2_rn(p[i], b2, c2); }
#pragma unroll
      for (int i = 0; i < CHAINS / 2; ++i) { float2 q = make_float2(a[2 * i], a[2 * i + 1]); q = __ffma2_rn(q, b2, c2); a[2 * i] = q.x; a[2 * i + 1] = q.y; }
    }
  }
  float s = 0;
  for (int i = 0; i < CHAINS; ++i) s += a[i] + (float)u[i];
  for (int i = 0; i < CHAINS / 2; ++i) s += p[i].x + p[i].y;
  if (s == 123.456f) out[0] = s;
}

template <int MODE>
void run(const char* name, double fmas_per_iter, int blocks_per_sm) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out;
  cudaMalloc(&out, 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int blocks = sms * blocks_per_sm;
  k<MODE><<<blocks, 256>>>(out, 0.999f, 1e-3f, 5);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MODE><<<blocks, 256>>>(out, 0.999f, 1e-3f, 5);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double fmas = (double)blocks * 256 * ITERS * fmas_per_iter;
  printf("{\"mode\": \"%s\", \"blocks_per_sm\": %d, \"ms\": %.3f, \"TFLOPs\": %.2f, \"fma_per_clk_per_sm_at_1965MHz\": %.1f}\n", name, blocks_per_sm, ms,
         2 * fmas / ms / 1e9, fmas / (ms * 1e-3) / sms / 1.965e9);
  cudaFree(out);
}

int main() {
  for (int bps : {4, 8}) {
    run<0>("FFMA r,r,r", 8, bps);
    run<1>("FFMA r,imm,imm", 8, bps);
    run<2>("FFMA2 pair,pair,pair", 8, bps);
    run<3>("FFMA2 pair,bcast,imm", 8, bps);
    run<4>("8 FFMA + 16 int", 8, bps);
    run<5>("4 FFMA2 + 16 int", 8, bps);
    run<6>("8 FFMA2 (16 fma)", 16, bps);
  }
  return 0;
}
