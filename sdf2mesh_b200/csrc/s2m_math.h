/* s2m_math.h -- pinned, bit-reproducible f32 math for SDF evaluation.
 *
 * The reference (WilstonOreo/sdf2mesh) leaves the meaning of sqrt/sin/cos/atan/asin/pow/log/...
 * to whatever Vulkan/Metal/DX12 driver compiler wgpu hands the shader to:
 * nothing under /root/reference pins them.  This header pins them for this engine.
 *
 * Every function below is built ONLY from operations that IEEE-754 defines exactly and that
 * x86-64 (SSE2 + FMA) and sm_100a implement identically when contraction is disabled:
 *   + - * / sqrt fma, comparisons, float<->int conversions of in-range values, integer ops.
 * The same text is compiled three ways:
 *   - NVRTC / nvcc for sm_100a (device: the product path)        --fmad=false
 *   - g++ for the CPU oracle under oracle/ (test infrastructure)  -ffp-contract=off -mfma
 *   - g++ for tests/ that check accuracy against libm (double)
 * so the device and the oracle agree bit-for-bit on every SDF value (NaN payloads excepted).
 *
 * Accuracy (measured by tests/test_math.py against libm in double): <= 2.4 ulp for
 * sin cos tan asin acos atan atan2 exp exp2 log log2 pow on their usual domains (<= 3 ulp for
 * sinh cosh tanh asinh acosh atanh); pow with an
 * integer exponent |n| <= 8 is a multiplication chain (<= 5 ulp).
 *
 * C99 / C++ / CUDA compatible.  No includes on the device (NVRTC has no libc headers).
 */
#ifndef S2M_MATH_H_
#define S2M_MATH_H_

#if defined(__CUDACC__) || defined(__CUDACC_RTC__)
#define S2M_HD __host__ __device__ __forceinline__
#else
#define S2M_HD static inline
#endif

#if !defined(__CUDACC_RTC__)
#include <math.h>
#include <stdint.h>
#endif

/* ------------------------------------------------------------------ bit casts, exact primitives */
S2M_HD int s2m_f2i(float x) {
#if defined(__CUDA_ARCH__)
  return __float_as_int(x);
#else
  int i; __builtin_memcpy(&i, &x, 4); return i;
#endif
}
S2M_HD float s2m_i2f(int i) {
#if defined(__CUDA_ARCH__)
  return __int_as_float(i);
#else
  float x; __builtin_memcpy(&x, &i, 4); return x;
#endif
}
S2M_HD float s2m_inf(void) { return s2m_i2f(0x7f800000); }
S2M_HD float s2m_nan(void) { return s2m_i2f(0x7fc00000); }
S2M_HD int s2m_isnan(float x) { return x != x; }

S2M_HD float s2m_fma(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
  return fmaf(a, b, c);
#else
  return __builtin_fmaf(a, b, c);
#endif
}
S2M_HD float s2m_sqrt(float x) { return sqrtf(x); } /* IEEE: sqrt.rn.f32 / sqrtss */
S2M_HD float s2m_abs(float x) { return s2m_i2f(s2m_f2i(x) & 0x7fffffff); }

/* min/max: WGSL leaves NaN and signed-zero behaviour open.  Pinned to IEEE-754-2019
 * minimumNumber/maximumNumber (= PTX min.f32/max.f32 = one FMNMX): a NaN operand is ignored,
 * and -0 < +0. */
S2M_HD float s2m_min(float a, float b) {
#if defined(__CUDA_ARCH__)
  return fminf(a, b);
#else
  if (a != a) return b;
  if (b != b) return a;
  if (a == b) return s2m_i2f(s2m_f2i(a) | s2m_f2i(b)); /* equal: pick -0 over +0 */
  return a < b ? a : b;
#endif
}
S2M_HD float s2m_max(float a, float b) {
#if defined(__CUDA_ARCH__)
  return fmaxf(a, b);
#else
  if (a != a) return b;
  if (b != b) return a;
  if (a == b) return s2m_i2f(s2m_f2i(a) & s2m_f2i(b)); /* equal: pick +0 over -0 */
  return a > b ? a : b;
#endif
}
/* WGSL spec: clamp(e,lo,hi) = min(max(e,lo),hi);  mix(a,b,t) = a*(1-t) + b*t */
S2M_HD float s2m_clamp(float x, float lo, float hi) { return s2m_min(s2m_max(x, lo), hi); }
S2M_HD float s2m_mix(float a, float b, float t) { return a * (1.0f - t) + b * t; }

S2M_HD float s2m_floor(float x) { return floorf(x); }
S2M_HD float s2m_ceil(float x) { return ceilf(x); }
S2M_HD float s2m_trunc(float x) { return truncf(x); }
S2M_HD float s2m_round(float x) { return rintf(x); } /* ties to even (WGSL round, GLSL roundEven) */
S2M_HD float s2m_fract(float x) { return x - floorf(x); }
S2M_HD float s2m_sign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : (x == 0.0f ? 0.0f : x)); }
S2M_HD float s2m_step(float edge, float x) { return edge <= x ? 1.0f : 0.0f; }
S2M_HD float s2m_smoothstep(float lo, float hi, float x) {
  float t = s2m_clamp((x - lo) / (hi - lo), 0.0f, 1.0f);
  return t * t * (3.0f - 2.0f * t);
}
/* WGSL % on floats: x - y*trunc(x/y);  GLSL mod(): x - y*floor(x/y) */
S2M_HD float s2m_fmod_trunc(float x, float y) { return x - y * truncf(x / y); }
S2M_HD float s2m_mod_floor(float x, float y) { return x - y * floorf(x / y); }
S2M_HD float s2m_inversesqrt(float x) { return 1.0f / sqrtf(x); }
S2M_HD float s2m_radians(float d) { return d * 1.745329238e-02f; }
S2M_HD float s2m_degrees(float r) { return r * 5.729578018e+01f; }

/* float -> int conversions saturate and map NaN to 0 (WGSL rule); never executes an
 * out-of-range cvt (x86 and PTX disagree on those). */
S2M_HD int s2m_f2int(float x) {
  if (x != x) return 0;
  if (x >= 2147483648.0f) return 2147483647;
  if (x <= -2147483648.0f) return (-2147483647 - 1);
  return (int)x;
}
S2M_HD unsigned s2m_f2uint(float x) {
  if (!(x > 0.0f)) return 0u;
  if (x >= 4294967296.0f) return 4294967295u;
  return (unsigned)x;
}

/* ------------------------------------------------------------------ trigonometric range reduction */
/* Fast path |x| <= 105615: 3-term Cody-Waite with FMA.  r = x - j*pi/2, *q = j mod 4. */
#if defined(__CUDACC__) || defined(__CUDACC_RTC__)
__host__ __device__ __noinline__ float s2m__trig_red_slow(float x, int* q);
#else
static float s2m__trig_red_slow(float x, int* q);
#endif

/* Fast reduction (Cody-Waite, 3 fma): valid for |x| <= 105615; straight-line, so that sin(x) and
 * cos(x) of one argument share it (and both polynomials) after CSE.  Larger arguments, inf and NaN
 * are patched afterwards by the callers through the out-of-line Payne-Hanek path.  The quadrant is
 * read from the low mantissa bits of the 1.5*2^23-biased product instead of a float->int convert. */
#define S2M__TRIG_FAST_MAX 105615.0f
S2M_HD float s2m__trig_red_fast(float x, int* q) {
  float jm = s2m_fma(x, 6.366197467e-01f, 12582912.0f);
  float j = jm - 12582912.0f; /* rint(x*2/pi) */
  float r = s2m_fma(j, -1.570796371e+00f, x);
  r = s2m_fma(j, 4.371138829e-08f, r);
  r = s2m_fma(j, 1.715124510e-15f, r);
  *q = s2m_f2i(jm) & 3;
  return r;
}
S2M_HD float s2m__trig_red(float x, int* q) {
  if (s2m_abs(x) > S2M__TRIG_FAST_MAX) return s2m__trig_red_slow(x, q);
  return s2m__trig_red_fast(x, q);
}

/* 2/pi = 0.A2F9836E 4E441529 FC2757D1 F534DDC0 DB629599 3C439041 FE5163AB ... (hex), one leading zero word */
#define S2M_TWO_OVER_PI_WORDS {0u, 0xa2f9836eu, 0x4e441529u, 0xfc2757d1u, 0xf534ddc0u, 0xdb629599u, 0x3c439041u, 0xfe5163abu, 0xdebbc561u}
#if defined(__CUDACC__) || defined(__CUDACC_RTC__)
static __constant__ unsigned s2m__two_over_pi_dev[9] = S2M_TWO_OVER_PI_WORDS;
#endif
#if !defined(__CUDA_ARCH__)
static const unsigned s2m__two_over_pi_host[9] = S2M_TWO_OVER_PI_WORDS;
#endif

/* Slow path: Payne-Hanek with integer arithmetic (inf/NaN -> NaN).  Out of line on the device:
 * it is cold, and inlining it at every sin/cos call site only bloats the SDF kernels. */
#if defined(__CUDACC__) || defined(__CUDACC_RTC__)
__host__ __device__ __noinline__
#else
static
#endif
float s2m__trig_red_slow(float x, int* q) {
#if defined(__CUDA_ARCH__)
  const unsigned* tw = s2m__two_over_pi_dev;
#else
  const unsigned* tw = s2m__two_over_pi_host;
#endif
  int ia = s2m_f2i(x) & 0x7fffffff;
  *q = 0;
  if (ia >= 0x7f800000) return s2m_nan();
  int e = (ia >> 23) - 127;                       /* >= 16 here */
  unsigned mant = (unsigned)((ia & 0x7fffff) | 0x800000);
  /* window of 96 bits of 2/pi starting at bit j0 = e - 24 (1-based), table has 32 leading zeros */
  int j0 = e - 24 + 32;                           /* >= 24 */
  int idx = (j0 - 1) >> 5, sh = (j0 - 1) & 31;
  unsigned w0, w1, w2;
  if (sh) {
    w0 = (tw[idx] << sh) | (tw[idx + 1] >> (32 - sh));
    w1 = (tw[idx + 1] << sh) | (tw[idx + 2] >> (32 - sh));
    w2 = (tw[idx + 2] << sh) | (tw[idx + 3] >> (32 - sh));
  } else {
    w0 = tw[idx]; w1 = tw[idx + 1]; w2 = tw[idx + 2];
  }
  unsigned long long p2 = (unsigned long long)mant * w2;
  unsigned long long p1 = (unsigned long long)mant * w1 + (p2 >> 32);
  unsigned long long p0 = (unsigned long long)mant * w0 + (p1 >> 32);
  unsigned long long h = (p0 << 32) | (p1 & 0xffffffffull); /* x*2/pi mod 4, scaled by 2^62 */
  int quad = (int)(h >> 62);
  long long f = (long long)(h & 0x3fffffffffffffffull);
  if (f >= (1ll << 61)) { f -= (1ll << 62); quad += 1; }
  double r = (double)f * 2.168404344971008868e-19 /* 2^-62 */ * 1.5707963267948966;
  float rf = (float)r;
  if (s2m_f2i(x) < 0) { rf = -rf; quad = -quad; }
  *q = quad & 3;
  return rf;
}

S2M_HD float s2m__sin_poly(float r) { /* |r| <= pi/4 */
  float s = r * r;
  float p = 2.717366897e-06f;
  p = s2m_fma(p, s, -1.983923285e-04f);
  p = s2m_fma(p, s, 8.333329111e-03f);
  p = s2m_fma(p, s, -1.666666716e-01f);
  return s2m_fma(p * s, r, r);
}
S2M_HD float s2m__cos_poly(float r) {
  float s = r * r;
  float p = -2.719862664e-07f;
  p = s2m_fma(p, s, 2.479937211e-05f);
  p = s2m_fma(p, s, -1.388888340e-03f);
  p = s2m_fma(p, s, 4.166666791e-02f);
  p = s2m_fma(p, s, -0.5f);
  return s2m_fma(p, s, 1.0f);
}
/* Branch-lean: both polynomials are always evaluated on the fast reduction and selected by the
 * quadrant; sin(x) and cos(x) of the same argument then share everything but the selects after CSE,
 * and lanes in different quadrants do not diverge.  Big arguments take the cold out-of-line path.
 * (Same values as reducing first and selecting before evaluating.) */
S2M_HD float s2m__sincos_sel(float r, int q) {
  const float sp = s2m__sin_poly(r), cp = s2m__cos_poly(r);
  const float v = (q & 1) ? cp : sp;
  return (q & 2) ? -v : v;
}
#if defined(__CUDACC__) || defined(__CUDACC_RTC__)
__host__ __device__ __noinline__
#else
static
#endif
float s2m__sincos_slow(float x, int dq) { /* |x| > S2M__TRIG_FAST_MAX, inf, NaN */
  int q; float r = s2m__trig_red_slow(x, &q);
  return s2m__sincos_sel(r, q + dq);
}
S2M_HD float s2m_sin(float x) {
  int q; float r = s2m__trig_red_fast(x, &q);
  float v = s2m__sincos_sel(r, q);
  if (s2m_abs(x) > S2M__TRIG_FAST_MAX) v = s2m__sincos_slow(x, 0);
  return v;
}
S2M_HD float s2m_cos(float x) {
  int q; float r = s2m__trig_red_fast(x, &q);
  float v = s2m__sincos_sel(r, q + 1);
  if (s2m_abs(x) > S2M__TRIG_FAST_MAX) v = s2m__sincos_slow(x, 1);
  return v;
}
/* sin and cos of one argument: one reduction, both polynomials, ONE cold big-argument test.  Same
 * values as s2m_sin(x) and s2m_cos(x); the front-end pairs the calls (optimize.cpp: pair_sin_cos). */
#if defined(__CUDACC__) || defined(__CUDACC_RTC__)
__host__ __device__ __noinline__
#else
static
#endif
void s2m__sincos_slow2(float x, float* s, float* c) {
  int q; float r = s2m__trig_red_slow(x, &q);
  *s = s2m__sincos_sel(r, q);
  *c = s2m__sincos_sel(r, q + 1);
}
S2M_HD void s2m_sincos(float x, float* s, float* c) {
  int q; float r = s2m__trig_red_fast(x, &q);
  const float sp = s2m__sin_poly(r), cp = s2m__cos_poly(r);
  const float vs = (q & 1) ? cp : sp, vc = (q & 1) ? sp : cp;
  *s = (q & 2) ? -vs : vs;
  *c = ((q + 1) & 2) ? -vc : vc;
  if (s2m_abs(x) > S2M__TRIG_FAST_MAX) s2m__sincos_slow2(x, s, c);
}
S2M_HD float s2m_tan(float x) {
  int q; float r = s2m__trig_red(x, &q);
  float s = r * r;
  float p = 4.486086778e-03f;
  p = s2m_fma(p, s, -1.291844965e-04f);
  p = s2m_fma(p, s, 1.100504305e-02f);
  p = s2m_fma(p, s, 2.121827193e-02f);
  p = s2m_fma(p, s, 5.407200754e-02f);
  p = s2m_fma(p, s, 1.333255917e-01f);
  p = s2m_fma(p, s, 3.333335221e-01f);
  float t = s2m_fma(p * s, r, r);
  return (q & 1) ? -1.0f / t : t;
}

/* ------------------------------------------------------------------ inverse trigonometric */
S2M_HD float s2m__atan_poly(float t) { /* 0 <= t <= 1 */
  float s = t * t;
  float p = -1.793615986e-03f;
  p = s2m_fma(p, s, 1.091458090e-02f);
  p = s2m_fma(p, s, -3.117780387e-02f);
  p = s2m_fma(p, s, 5.795755610e-02f);
  p = s2m_fma(p, s, -8.403448015e-02f);
  p = s2m_fma(p, s, 1.095218509e-01f);
  p = s2m_fma(p, s, -1.426424086e-01f);
  p = s2m_fma(p, s, 1.999854892e-01f);
  p = s2m_fma(p, s, -3.333329856e-01f);
  return s2m_fma(p * s, t, t);
}
S2M_HD float s2m_atan(float x) {
  const float a = s2m_abs(x);
  const bool big = a > 1.0f;                     /* NaN: false, flows through the polynomial */
  const float t = big ? 1.0f / a : a;            /* 1/inf = 0 -> pi/2 */
  float r = s2m__atan_poly(t);
  if (big) r = s2m_fma(1.0f, 1.570796371e+00f, -r) + (-4.371138829e-08f);
  return s2m_i2f(s2m_f2i(r) | (s2m_f2i(x) & (int)0x80000000));
}
S2M_HD float s2m_atan2(float y, float x) {
  float ax = s2m_abs(x), ay = s2m_abs(y);
  float r;
  if (s2m_isnan(x) || s2m_isnan(y)) return s2m_nan();
  if (ay == 0.0f) {
    r = (s2m_f2i(x) < 0) ? 3.141592741e+00f : 0.0f;
  } else if (ax == ay) { /* includes inf/inf */
    r = (s2m_f2i(x) < 0) ? 2.356194496e+00f : 7.853981853e-01f;
  } else {
    float mx = ax > ay ? ax : ay, mn = ax > ay ? ay : ax;
    r = s2m__atan_poly(mn / mx);
    if (ay > ax) r = s2m_fma(1.0f, 1.570796371e+00f, -r) + (-4.371138829e-08f);
    if (s2m_f2i(x) < 0) r = (3.141592741e+00f - r) + (-8.742277657e-08f);
  }
  return s2m_i2f(s2m_f2i(r) | (s2m_f2i(y) & (int)0x80000000));
}
S2M_HD float s2m__asin_poly(float x, float s) { /* x + x*s*P(s) */
  float p = 3.751632944e-02f;
  p = s2m_fma(p, s, 1.443869714e-02f);
  p = s2m_fma(p, s, 3.180769086e-02f);
  p = s2m_fma(p, s, 4.451695830e-02f);
  p = s2m_fma(p, s, 7.500503957e-02f);
  p = s2m_fma(p, s, 1.666665971e-01f);
  return s2m_fma(p * s, x, x);
}
S2M_HD float s2m_asin(float x) {
  const float a = s2m_abs(x);
  const bool big = a > 0.5f;   /* asin(a) = pi/2 - 2*asin(sqrt((1-a)/2)); a > 1 -> NaN via sqrt */
  const float z = big ? s2m_fma(a, -0.5f, 0.5f) : a * a;
  const float y = big ? s2m_sqrt(z) : a;
  float r = s2m__asin_poly(y, z);
  if (big) r = s2m_fma(r, -2.0f, 1.570796371e+00f) + (-4.371138829e-08f);
  return s2m_i2f(s2m_f2i(r) | (s2m_f2i(x) & (int)0x80000000));
}
S2M_HD float s2m_acos(float x) {
  float a = s2m_abs(x);
  if (a > 0.5f) {
    float z = s2m_fma(a, -0.5f, 0.5f);
    float y = s2m_sqrt(z);
    float r = 2.0f * s2m__asin_poly(y, z);       /* acos(|x|) */
    return (s2m_f2i(x) < 0) ? (3.141592741e+00f - r) + (-8.742277657e-08f) : r;
  }
  float r = s2m__asin_poly(x, x * x);
  return (1.570796371e+00f - r) + (-4.371138829e-08f);
}

/* ------------------------------------------------------------------ exp / log / pow */
S2M_HD float s2m__scale2(float r, int i) { /* r * 2^i, i in [-252, 254] */
  int h = i >> 1;
  float s1 = s2m_i2f((127 + h) << 23);
  float s2 = s2m_i2f((127 + (i - h)) << 23);
  return (r * s1) * s2;
}
S2M_HD float s2m__exp_parts(float a, int* i) { /* exp(a) = ret * 2^i, |a| <= ~104 */
  float j = s2m_fma(a, 1.442695022e+00f, 12582912.0f) - 12582912.0f;
  float f = s2m_fma(j, -6.931471825e-01f, a);
  f = s2m_fma(j, 1.904654212e-09f, f);
  float p = 1.978926593e-04f;
  p = s2m_fma(p, f, 1.394575229e-03f);
  p = s2m_fma(p, f, 8.333504200e-03f);
  p = s2m_fma(p, f, 4.166628048e-02f);
  p = s2m_fma(p, f, 1.666666567e-01f);
  p = s2m_fma(p, f, 0.5f);
  *i = (int)j;
  return s2m_fma(p * f, f, f) + 1.0f;
}
S2M_HD float s2m__exp_core(float a) {
  int i; float r = s2m__exp_parts(a, &i);
  return s2m__scale2(r, i);
}
S2M_HD float s2m_exp(float a) {
  if (a != a) return a + a;
  if (a > 88.7228394f) return s2m_inf();
  if (a < -103.98f) return 0.0f;
  return s2m__exp_core(a);
}
S2M_HD float s2m_exp2(float a) {
  if (a != a) return a + a;
  if (a >= 128.0f) return s2m_inf();
  if (a < -150.0f) return 0.0f;
  float j = (a + 12582912.0f) - 12582912.0f; /* rint(a) */
  float f = a - j;                           /* exact, |f| <= 0.5 */
  float p = 1.519615398e-05f;
  p = s2m_fma(p, f, 1.546675630e-04f);
  p = s2m_fma(p, f, 1.333393506e-03f);
  p = s2m_fma(p, f, 9.618038312e-03f);
  p = s2m_fma(p, f, 5.550410226e-02f);
  p = s2m_fma(p, f, 2.402265072e-01f);
  p = s2m_fma(p, f, 6.931471825e-01f);
  float r = s2m_fma(p, f, 1.0f);
  return s2m__scale2(r, (int)j);
}

S2M_HD float s2m__log1p_poly(float f) { /* log1p(f) - f + f*f/2 = f^3 L(f),  |f| <= 1/3 */
  float p = -1.289160103e-01f;
  p = s2m_fma(p, f, 1.398446709e-01f);
  p = s2m_fma(p, f, -1.218427792e-01f);
  p = s2m_fma(p, f, 1.400586218e-01f);
  p = s2m_fma(p, f, -1.668048650e-01f);
  p = s2m_fma(p, f, 2.001040578e-01f);
  p = s2m_fma(p, f, -2.499979734e-01f);
  p = s2m_fma(p, f, 3.333321512e-01f);
  return p;
}
/* log / log2 of a normal positive finite float (bias = exponent correction of a pre-scaled denormal) */
S2M_HD float s2m__log_norm(float a, float bias) {
  const int ia = s2m_f2i(a);
  const int e = (ia - 0x3f2aaaab) & (int)0xff800000;   /* m = a / 2^i in [2/3, 4/3) */
  const float i = s2m_fma((float)e, 1.192092896e-07f, bias);
  const float f = s2m_i2f(ia - e) - 1.0f;
  const float s = f * f;
  float r = s2m_fma(s2m__log1p_poly(f) * f, s, i * -1.904654212e-09f);
  r = s2m_fma(-0.5f, s, r);
  r = r + f;
  return s2m_fma(i, 6.931471825e-01f, r);
}
S2M_HD float s2m__log2_norm(float a, float bias) {
  const int ia = s2m_f2i(a);
  const int e = (ia - 0x3f2aaaab) & (int)0xff800000;
  const float i = s2m_fma((float)e, 1.192092896e-07f, bias);
  const float f = s2m_i2f(ia - e) - 1.0f;
  const float s = f * f;
  float r = s2m_fma(s2m__log1p_poly(f) * f, s, -0.5f * s);   /* log1p(f) - f */
  /* (f + r) * log2(e) with log2(e) = hi + lo */
  float t = s2m_fma(f, 1.925963034e-08f, r * 1.442695022e+00f);
  t = s2m_fma(f, 1.442695022e+00f, t);
  return t + i;
}
/* zero, negative, denormal, inf, NaN: one cold out-of-line path behind a single range test */
#if defined(__CUDACC__) || defined(__CUDACC_RTC__)
__host__ __device__ __noinline__
#else
static
#endif
float s2m__log_special(float a, int base2) {
  const int ia = s2m_f2i(a);
  if ((ia & 0x7fffffff) == 0) return -s2m_inf();
  if (ia < 0) return s2m_nan();
  if (ia >= 0x7f800000) return a + a; /* +inf, NaN */
  a = a * 8388608.0f;                 /* denormal */
  return base2 ? s2m__log2_norm(a, -23.0f) : s2m__log_norm(a, -23.0f);
}
#define S2M__LOG_IS_SPECIAL(a) ((unsigned)(s2m_f2i(a) - 0x00800000) >= 0x7f000000u)
S2M_HD float s2m_log(float a) {
  if (S2M__LOG_IS_SPECIAL(a)) return s2m__log_special(a, 0);
  return s2m__log_norm(a, 0.0f);
}
S2M_HD float s2m_log2(float a) {
  if (S2M__LOG_IS_SPECIAL(a)) return s2m__log_special(a, 1);
  return s2m__log2_norm(a, 0.0f);
}

/* log(a) as a double-float hi:lo (a finite, positive); relative error ~1e-9. */
S2M_HD void s2m__log_ext(float a, float* hi, float* lo) {
  float bias = 0.0f;
  int ia = s2m_f2i(a);
  if (ia < 0x00800000) { a = a * 8388608.0f; bias = -23.0f; ia = s2m_f2i(a); }
  int e = (ia - 0x3f3504f3) & (int)0xff800000;      /* m in [sqrt(1/2), sqrt(2)) */
  float m = s2m_i2f(ia - e);
  float i = s2m_fma((float)e, 1.192092896e-07f, bias);
  /* q = (m-1)/(m+1) as qhi:qlo */
  float p = m + 1.0f;
  float d = m - 1.0f;
  float rcp = 1.0f / p;
  float qhi = rcp * d;
  float qlo = rcp * s2m_fma(qhi, -d, s2m_fma(qhi, -2.0f, d));
  /* atanh(q) = q + q^3 H(q^2) */
  float s = qhi * qhi;
  float r = 1.179095805e-01f;
  r = s2m_fma(r, s, 1.426866204e-01f);
  r = s2m_fma(r, s, 2.000016719e-01f);
  r = s2m_fma(r, s, 3.333333135e-01f);
  float t = s2m_fma(qhi, qlo + qlo, s2m_fma(qhi, qhi, -s));       /* s:t = q^2 */
  float c = s * qhi;
  t = s2m_fma(s, qlo, s2m_fma(t, qhi, s2m_fma(s, qhi, -c)));      /* c:t = q^3 */
  s = s2m_fma(r, c, s2m_fma(r, t, qlo));
  r = 2.0f * qhi;
  /* log(a) = 2*atanh(q) + i*log(2) */
  t = s2m_fma(6.931471825e-01f, i, r);
  c = s2m_fma(-6.931471825e-01f, i, t);
  s = s2m_fma(-1.904654212e-09f, i, s2m_fma(2.0f, s, r - c));
  c = t + s;
  *hi = c;
  *lo = (t - c) + s;
}
S2M_HD float s2m__pow_pos(float a, float b) { /* a finite > 0, b finite */
  float lhi, llo;
  s2m__log_ext(a, &lhi, &llo);
  float thi = lhi * b;
  if (thi > 88.7228394f) return s2m_inf();
  if (thi < -103.98f) return 0.0f;
  float tlo = s2m_fma(lhi, b, -thi);
  tlo = s2m_fma(llo, b, tlo);
  int i; float r = s2m__exp_parts(thi, &i);
  r = s2m_fma(tlo, r, r);
  return s2m__scale2(r, i);
}
/* C99 powf semantics (a superset of WGSL/GLSL pow, which leave a < 0 undefined). */
S2M_HD float s2m_pow(float a, float b) {
  const float ab = s2m_abs(b);
  /* Integer exponents up to 8 in magnitude: exponentiation by squaring (<= 5 multiplications,
   * <= 5 ulp measured; handles signs, zeros, infinities, a == 1, b == 0 and NaN a by itself, so it
   * comes before the special-case tests).  This is the strength reduction shader compilers apply
   * to pow(x, 8.0); when b is a compile-time constant it folds to the bare multiplication chain. */
  if (ab <= 8.0f && truncf(b) == b) {
    int n = (int)ab;
    float r = 1.0f, p = a;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 4; ++i) {
      if (n & (1 << i)) r = r * p;
      p = p * p;
    }
    return b < 0.0f ? 1.0f / r : r;
  }
  if (a == 1.0f) return 1.0f;
  if (s2m_isnan(a) || s2m_isnan(b)) return s2m_nan();
  const float aa = s2m_abs(a);
  int b_int = (ab >= 8388608.0f) || (truncf(b) == b);
  int b_odd = b_int && (ab < 16777216.0f) && ((((int)truncf(ab)) & 1) != 0) && (ab >= 1.0f);
  if (ab == s2m_inf()) {
    if (aa == 1.0f) return 1.0f;
    return ((aa > 1.0f) == (b > 0.0f)) ? s2m_inf() : 0.0f;
  }
  float sgn = (s2m_f2i(a) < 0 && b_odd) ? -1.0f : 1.0f;
  if (aa == 0.0f) return (b > 0.0f) ? sgn * 0.0f : sgn * s2m_inf();
  if (aa == s2m_inf()) return (b > 0.0f) ? sgn * s2m_inf() : sgn * 0.0f;
  if (s2m_f2i(a) < 0 && !b_int) return s2m_nan();
  return sgn * s2m__pow_pos(aa, b);
}

/* ------------------------------------------------------------------ hyperbolic (from exp) */
S2M_HD float s2m_sinh(float x) {
  float a = s2m_abs(x), r;
  if (a < 1.0f) {
    float s = a * a;
    float p = 2.816951110e-06f;
    p = s2m_fma(p, s, 1.983615948e-04f);
    p = s2m_fma(p, s, 8.333349600e-03f);
    p = s2m_fma(p, s, 1.666666716e-01f);
    r = s2m_fma(p * s, a, a);
  } else if (a > 89.5f) {
    r = (a != a) ? a : s2m_inf();
  } else {
    if (a < 88.0f) { float e = s2m__exp_core(a); r = 0.5f * e - 0.5f / e; }
    else r = s2m__exp_core(a - 64.0f) * 3.117574540e+27f; /* e^64 / 2; a-64 is exact */
  }
  return s2m_i2f(s2m_f2i(r) | (s2m_f2i(x) & (int)0x80000000));
}
S2M_HD float s2m_cosh(float x) {
  float a = s2m_abs(x);
  if (a != a) return a;
  if (a > 89.5f) return s2m_inf();
  if (a < 88.0f) { float e = s2m__exp_core(a); return 0.5f * e + 0.5f / e; }
  return s2m__exp_core(a - 64.0f) * 3.117574540e+27f;
}
S2M_HD float s2m_tanh(float x) {
  float a = s2m_abs(x), r;
  if (a != a) return x;
  if (a < 0.55f) {
    float s = a * a;
    float p = 2.395824064e-03f;
    p = s2m_fma(p, s, -8.413959295e-03f);
    p = s2m_fma(p, s, 2.178282663e-02f);
    p = s2m_fma(p, s, -5.395964906e-02f);
    p = s2m_fma(p, s, 1.333329380e-01f);
    p = s2m_fma(p, s, -3.333333135e-01f);
    r = s2m_fma(p * s, a, a);
  } else if (a > 9.1f) {
    r = 1.0f;
  } else {
    float e = s2m__exp_core(2.0f * a);
    r = 1.0f - 2.0f / (e + 1.0f);
  }
  return s2m_i2f(s2m_f2i(r) | (s2m_f2i(x) & (int)0x80000000));
}

/* ------------------------------------------------------------------ inverse hyperbolic (from log)
 * log1p(u) for u >= 0 through log: w = fl(1 + u) carries a rounding error that the second term removes
 * (log(1+u) = log(w) + log((1+u)/w) ~ log(w) - ((w - 1) - u) / w; w - 1 is exact). */
S2M_HD float s2m__log1p_pos(float u) {
  float w = 1.0f + u;
  if (w == 1.0f) return u;
  if (w == s2m_inf()) return w;
  return s2m_log(w) - ((w - 1.0f) - u) / w;
}
S2M_HD float s2m_asinh(float x) {   /* log(a + sqrt(a^2 + 1)) = log1p(a + a^2 / (1 + sqrt(a^2 + 1))) */
  float a = s2m_abs(x), r;
  if (a != a) return x;
  if (a > 1.0e9f) r = s2m_log(a) + 6.931471825e-01f;            /* a^2 would overflow; sqrt(a^2 + 1) = a to 1e-18 */
  else { float s = a * a; r = s2m__log1p_pos(a + s / (1.0f + s2m_sqrt(s + 1.0f))); }
  return s2m_i2f(s2m_f2i(r) | (s2m_f2i(x) & (int)0x80000000));
}
S2M_HD float s2m_acosh(float x) {   /* log(x + sqrt(x^2 - 1)) = log1p(d + sqrt(d (d + 2))), d = x - 1 */
  if (!(x >= 1.0f)) return s2m_nan();
  if (x > 1.0e9f) return s2m_log(x) + 6.931471825e-01f;
  float d = x - 1.0f;
  return s2m__log1p_pos(d + s2m_sqrt(d * (d + 2.0f)));
}
S2M_HD float s2m_atanh(float x) {   /* log((1 + a) / (1 - a)) / 2 = log1p(2a / (1 - a)) / 2 */
  float a = s2m_abs(x), r;
  if (!(a <= 1.0f)) return s2m_nan();
  r = 0.5f * s2m__log1p_pos((a + a) / (1.0f - a));               /* a = 1: 2 / 0 = inf */
  return s2m_i2f(s2m_f2i(r) | (s2m_f2i(x) & (int)0x80000000));
}

#endif /* S2M_MATH_H_ */
