/* s2m_vec.h -- vector types and component-wise builtins for SDF code lowered from WGSL / GLSL.
 *
 * The emitter (frontend/emit_cuda.cpp) lowers naga-shaped IR to C++ that uses these types; the
 * same text compiles under NVRTC (device, --fmad=false) and g++ (tests, -ffp-contract=off), so
 * every expression is evaluated op-by-op in IEEE f32 in the order the shader wrote it.
 *
 * Pinned semantics for what WGSL/GLSL leave to the driver:
 *   dot = products summed left to right; length = sqrt(dot(v,v)); normalize = v / length(v);
 *   distance(a,b) = length(a-b); cross per the usual formula (a.y*b.z - a.z*b.y, ...);
 *   reflect(i,n) = i - 2*dot(n,i)*n.
 */
#ifndef S2M_VEC_H_
#define S2M_VEC_H_
#include "s2m_math.h"

namespace s2m {

struct vec2 { float x, y; };
struct vec3 { float x, y, z; };
struct vec4 { float x, y, z, w; };
struct ivec2 { int x, y; };
struct ivec3 { int x, y, z; };
struct ivec4 { int x, y, z, w; };
struct uvec2 { unsigned x, y; };
struct uvec3 { unsigned x, y, z; };
struct uvec4 { unsigned x, y, z, w; };
struct bvec2 { bool x, y; };
struct bvec3 { bool x, y, z; };
struct bvec4 { bool x, y, z, w; };

S2M_HD vec2 mk2(float x, float y) { vec2 v; v.x = x; v.y = y; return v; }
S2M_HD vec3 mk3(float x, float y, float z) { vec3 v; v.x = x; v.y = y; v.z = z; return v; }
S2M_HD vec4 mk4(float x, float y, float z, float w) { vec4 v; v.x = x; v.y = y; v.z = z; v.w = w; return v; }
S2M_HD ivec2 mki2(int x, int y) { ivec2 v; v.x = x; v.y = y; return v; }
S2M_HD ivec3 mki3(int x, int y, int z) { ivec3 v; v.x = x; v.y = y; v.z = z; return v; }
S2M_HD ivec4 mki4(int x, int y, int z, int w) { ivec4 v; v.x = x; v.y = y; v.z = z; v.w = w; return v; }
S2M_HD uvec2 mku2(unsigned x, unsigned y) { uvec2 v; v.x = x; v.y = y; return v; }
S2M_HD uvec3 mku3(unsigned x, unsigned y, unsigned z) { uvec3 v; v.x = x; v.y = y; v.z = z; return v; }
S2M_HD uvec4 mku4(unsigned x, unsigned y, unsigned z, unsigned w) { uvec4 v; v.x = x; v.y = y; v.z = z; v.w = w; return v; }
S2M_HD uvec3 mku3(const uvec2& a, unsigned b) { return mku3(a.x, a.y, b); }
S2M_HD uvec3 mku3(unsigned a, const uvec2& b) { return mku3(a, b.x, b.y); }
S2M_HD uvec4 mku4(const uvec3& a, unsigned b) { return mku4(a.x, a.y, a.z, b); }
S2M_HD uvec4 mku4(const uvec2& a, const uvec2& b) { return mku4(a.x, a.y, b.x, b.y); }
S2M_HD uvec2 splatu2(unsigned a) { return mku2(a, a); }
S2M_HD uvec3 splatu3(unsigned a) { return mku3(a, a, a); }
S2M_HD uvec4 splatu4(unsigned a) { return mku4(a, a, a, a); }
S2M_HD bvec2 mkb2(bool x, bool y) { bvec2 v; v.x = x; v.y = y; return v; }
S2M_HD bvec3 mkb3(bool x, bool y, bool z) { bvec3 v; v.x = x; v.y = y; v.z = z; return v; }
S2M_HD bvec4 mkb4(bool x, bool y, bool z, bool w) { bvec4 v; v.x = x; v.y = y; v.z = z; v.w = w; return v; }


/* mixed-argument constructors: vec3(v2, s), vec4(v3, s), vec4(v2, v2), ... and splats */
S2M_HD vec3 mk3(const vec2& a, float b) { return mk3(a.x, a.y, b); }
S2M_HD vec3 mk3(float a, const vec2& b) { return mk3(a, b.x, b.y); }
S2M_HD vec4 mk4(const vec3& a, float b) { return mk4(a.x, a.y, a.z, b); }
S2M_HD vec4 mk4(float a, const vec3& b) { return mk4(a, b.x, b.y, b.z); }
S2M_HD vec4 mk4(const vec2& a, const vec2& b) { return mk4(a.x, a.y, b.x, b.y); }
S2M_HD vec4 mk4(const vec2& a, float b, float c) { return mk4(a.x, a.y, b, c); }
S2M_HD vec4 mk4(float a, const vec2& b, float c) { return mk4(a, b.x, b.y, c); }
S2M_HD vec4 mk4(float a, float b, const vec2& c) { return mk4(a, b, c.x, c.y); }
S2M_HD ivec3 mki3(const ivec2& a, int b) { return mki3(a.x, a.y, b); }
S2M_HD ivec3 mki3(int a, const ivec2& b) { return mki3(a, b.x, b.y); }
S2M_HD ivec4 mki4(const ivec3& a, int b) { return mki4(a.x, a.y, a.z, b); }
S2M_HD ivec4 mki4(const ivec2& a, const ivec2& b) { return mki4(a.x, a.y, b.x, b.y); }
S2M_HD vec2 splat2(float a) { return mk2(a, a); }
S2M_HD vec3 splat3(float a) { return mk3(a, a, a); }
S2M_HD vec4 splat4(float a) { return mk4(a, a, a, a); }
S2M_HD ivec2 splati2(int a) { return mki2(a, a); }
S2M_HD ivec3 splati3(int a) { return mki3(a, a, a); }
S2M_HD ivec4 splati4(int a) { return mki4(a, a, a, a); }
S2M_HD bvec2 splatb2(bool a) { return mkb2(a, a); }
S2M_HD bvec3 splatb3(bool a) { return mkb3(a, a, a); }
S2M_HD bvec4 splatb4(bool a) { return mkb4(a, a, a, a); }

/* component access by index (swizzles are lowered to these) */
S2M_HD float cget(const vec2& v, int i) { return i == 0 ? v.x : v.y; }
S2M_HD float cget(const vec3& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }
S2M_HD float cget(const vec4& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }
S2M_HD int cget(const ivec2& v, int i) { return i == 0 ? v.x : v.y; }
S2M_HD int cget(const ivec3& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }
S2M_HD int cget(const ivec4& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }
S2M_HD bool cget(const bvec2& v, int i) { return i == 0 ? v.x : v.y; }
S2M_HD bool cget(const bvec3& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }
S2M_HD bool cget(const bvec4& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }
S2M_HD void cset(vec2& v, int i, float f) { if (i == 0) v.x = f; else v.y = f; }
S2M_HD void cset(vec3& v, int i, float f) { if (i == 0) v.x = f; else if (i == 1) v.y = f; else v.z = f; }
S2M_HD void cset(vec4& v, int i, float f) { if (i == 0) v.x = f; else if (i == 1) v.y = f; else if (i == 2) v.z = f; else v.w = f; }
S2M_HD void cset(ivec2& v, int i, int f) { if (i == 0) v.x = f; else v.y = f; }
S2M_HD void cset(ivec3& v, int i, int f) { if (i == 0) v.x = f; else if (i == 1) v.y = f; else v.z = f; }
S2M_HD void cset(ivec4& v, int i, int f) { if (i == 0) v.x = f; else if (i == 1) v.y = f; else if (i == 2) v.z = f; else v.w = f; }


S2M_HD unsigned cget(const uvec2& v, int i) { return i == 0 ? v.x : v.y; }
S2M_HD unsigned cget(const uvec3& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }
S2M_HD unsigned cget(const uvec4& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }

/* multi-component swizzle reads: swz3(v, 0, 2, 1) == v.xzy */
S2M_HD vec2 swz2(const vec2& v, int a, int b) { return mk2(cget(v, a), cget(v, b)); }
S2M_HD vec2 swz2(const vec3& v, int a, int b) { return mk2(cget(v, a), cget(v, b)); }
S2M_HD vec2 swz2(const vec4& v, int a, int b) { return mk2(cget(v, a), cget(v, b)); }
S2M_HD vec3 swz3(const vec2& v, int a, int b, int c) { return mk3(cget(v, a), cget(v, b), cget(v, c)); }
S2M_HD vec3 swz3(const vec3& v, int a, int b, int c) { return mk3(cget(v, a), cget(v, b), cget(v, c)); }
S2M_HD vec3 swz3(const vec4& v, int a, int b, int c) { return mk3(cget(v, a), cget(v, b), cget(v, c)); }
S2M_HD vec4 swz4(const vec2& v, int a, int b, int c, int d) { return mk4(cget(v, a), cget(v, b), cget(v, c), cget(v, d)); }
S2M_HD vec4 swz4(const vec3& v, int a, int b, int c, int d) { return mk4(cget(v, a), cget(v, b), cget(v, c), cget(v, d)); }
S2M_HD vec4 swz4(const vec4& v, int a, int b, int c, int d) { return mk4(cget(v, a), cget(v, b), cget(v, c), cget(v, d)); }
S2M_HD ivec2 swz2(const ivec2& v, int a, int b) { return mki2(cget(v, a), cget(v, b)); }
S2M_HD ivec2 swz2(const ivec3& v, int a, int b) { return mki2(cget(v, a), cget(v, b)); }
S2M_HD ivec2 swz2(const ivec4& v, int a, int b) { return mki2(cget(v, a), cget(v, b)); }
S2M_HD ivec3 swz3(const ivec3& v, int a, int b, int c) { return mki3(cget(v, a), cget(v, b), cget(v, c)); }
S2M_HD ivec3 swz3(const ivec4& v, int a, int b, int c) { return mki3(cget(v, a), cget(v, b), cget(v, c)); }

/* ---- arithmetic: vec op vec, vec op scalar, scalar op vec, unary minus */
#define S2M_VEC_BINOP(OP)                                                                             \
  S2M_HD vec2 operator OP(const vec2& a, const vec2& b) { return mk2(a.x OP b.x, a.y OP b.y); }       \
  S2M_HD vec3 operator OP(const vec3& a, const vec3& b) { return mk3(a.x OP b.x, a.y OP b.y, a.z OP b.z); } \
  S2M_HD vec4 operator OP(const vec4& a, const vec4& b) { return mk4(a.x OP b.x, a.y OP b.y, a.z OP b.z, a.w OP b.w); } \
  S2M_HD vec2 operator OP(const vec2& a, float b) { return mk2(a.x OP b, a.y OP b); }                 \
  S2M_HD vec3 operator OP(const vec3& a, float b) { return mk3(a.x OP b, a.y OP b, a.z OP b); }       \
  S2M_HD vec4 operator OP(const vec4& a, float b) { return mk4(a.x OP b, a.y OP b, a.z OP b, a.w OP b); } \
  S2M_HD vec2 operator OP(float a, const vec2& b) { return mk2(a OP b.x, a OP b.y); }                 \
  S2M_HD vec3 operator OP(float a, const vec3& b) { return mk3(a OP b.x, a OP b.y, a OP b.z); }       \
  S2M_HD vec4 operator OP(float a, const vec4& b) { return mk4(a OP b.x, a OP b.y, a OP b.z, a OP b.w); }
S2M_VEC_BINOP(+)
S2M_VEC_BINOP(-)
S2M_VEC_BINOP(*)
S2M_VEC_BINOP(/)
#undef S2M_VEC_BINOP
S2M_HD vec2 operator-(const vec2& a) { return mk2(-a.x, -a.y); }
S2M_HD vec3 operator-(const vec3& a) { return mk3(-a.x, -a.y, -a.z); }
S2M_HD vec4 operator-(const vec4& a) { return mk4(-a.x, -a.y, -a.z, -a.w); }

/* ---- integer scalars and vectors.  Pinned to WGSL's rules where C++ would be undefined:
 * x / 0 = x, x % 0 = 0, INT_MIN / -1 = INT_MIN, INT_MIN % -1 = 0; shift counts are taken mod 32;
 * signed overflow wraps (computed in unsigned). */
S2M_HD int i_add(int a, int b) { return (int)((unsigned)a + (unsigned)b); }
S2M_HD int i_sub(int a, int b) { return (int)((unsigned)a - (unsigned)b); }
S2M_HD int i_mul(int a, int b) { return (int)((unsigned)a * (unsigned)b); }
S2M_HD int i_div(int a, int b) { return (b == 0 || (a == (-2147483647 - 1) && b == -1)) ? a : a / b; }
S2M_HD int i_rem(int a, int b) { return (b == 0 || (a == (-2147483647 - 1) && b == -1)) ? 0 : a % b; }
S2M_HD int i_shl(int a, unsigned b) { return (int)((unsigned)a << (b & 31u)); }
S2M_HD int i_shr(int a, unsigned b) { return a >> (b & 31u); }
S2M_HD int i_shl(int a, int b) { return i_shl(a, (unsigned)b); }
S2M_HD int i_shr(int a, int b) { return i_shr(a, (unsigned)b); }
S2M_HD int i_and(int a, int b) { return a & b; }
S2M_HD int i_or(int a, int b) { return a | b; }
S2M_HD int i_xor(int a, int b) { return a ^ b; }
S2M_HD int i_neg(int a) { return (int)(0u - (unsigned)a); }
S2M_HD int i_not(int a) { return ~a; }
S2M_HD unsigned i_add(unsigned a, unsigned b) { return a + b; }
S2M_HD unsigned i_sub(unsigned a, unsigned b) { return a - b; }
S2M_HD unsigned i_mul(unsigned a, unsigned b) { return a * b; }
S2M_HD unsigned i_div(unsigned a, unsigned b) { return b == 0u ? a : a / b; }
S2M_HD unsigned i_rem(unsigned a, unsigned b) { return b == 0u ? 0u : a % b; }
S2M_HD unsigned i_shl(unsigned a, unsigned b) { return a << (b & 31u); }
S2M_HD unsigned i_shr(unsigned a, unsigned b) { return a >> (b & 31u); }
S2M_HD unsigned i_shl(unsigned a, int b) { return i_shl(a, (unsigned)b); }
S2M_HD unsigned i_shr(unsigned a, int b) { return i_shr(a, (unsigned)b); }
S2M_HD unsigned i_and(unsigned a, unsigned b) { return a & b; }
S2M_HD unsigned i_or(unsigned a, unsigned b) { return a | b; }
S2M_HD unsigned i_xor(unsigned a, unsigned b) { return a ^ b; }
S2M_HD unsigned i_neg(unsigned a) { return 0u - a; }
S2M_HD unsigned i_not(unsigned a) { return ~a; }
#define S2M_INT_VEC2(NAME, V2, V3, V4, T, M2, M3, M4)                                                  \
  S2M_HD V2 NAME(const V2& a, const V2& b) { return M2(NAME(a.x, b.x), NAME(a.y, b.y)); }              \
  S2M_HD V3 NAME(const V3& a, const V3& b) { return M3(NAME(a.x, b.x), NAME(a.y, b.y), NAME(a.z, b.z)); } \
  S2M_HD V4 NAME(const V4& a, const V4& b) { return M4(NAME(a.x, b.x), NAME(a.y, b.y), NAME(a.z, b.z), NAME(a.w, b.w)); } \
  S2M_HD V2 NAME(const V2& a, T b) { return M2(NAME(a.x, b), NAME(a.y, b)); }                          \
  S2M_HD V3 NAME(const V3& a, T b) { return M3(NAME(a.x, b), NAME(a.y, b), NAME(a.z, b)); }            \
  S2M_HD V4 NAME(const V4& a, T b) { return M4(NAME(a.x, b), NAME(a.y, b), NAME(a.z, b), NAME(a.w, b)); } \
  S2M_HD V2 NAME(T a, const V2& b) { return M2(NAME(a, b.x), NAME(a, b.y)); }                          \
  S2M_HD V3 NAME(T a, const V3& b) { return M3(NAME(a, b.x), NAME(a, b.y), NAME(a, b.z)); }            \
  S2M_HD V4 NAME(T a, const V4& b) { return M4(NAME(a, b.x), NAME(a, b.y), NAME(a, b.z), NAME(a, b.w)); }
#define S2M_INT_VEC1(NAME, V2, V3, V4, M2, M3, M4)                                                     \
  S2M_HD V2 NAME(const V2& a) { return M2(NAME(a.x), NAME(a.y)); }                                     \
  S2M_HD V3 NAME(const V3& a) { return M3(NAME(a.x), NAME(a.y), NAME(a.z)); }                          \
  S2M_HD V4 NAME(const V4& a) { return M4(NAME(a.x), NAME(a.y), NAME(a.z), NAME(a.w)); }
#define S2M_INT_FAMILY(V2, V3, V4, T, M2, M3, M4)                                                      \
  S2M_INT_VEC2(i_add, V2, V3, V4, T, M2, M3, M4) S2M_INT_VEC2(i_sub, V2, V3, V4, T, M2, M3, M4)        \
  S2M_INT_VEC2(i_mul, V2, V3, V4, T, M2, M3, M4) S2M_INT_VEC2(i_div, V2, V3, V4, T, M2, M3, M4)        \
  S2M_INT_VEC2(i_rem, V2, V3, V4, T, M2, M3, M4) S2M_INT_VEC2(i_and, V2, V3, V4, T, M2, M3, M4)        \
  S2M_INT_VEC2(i_or, V2, V3, V4, T, M2, M3, M4) S2M_INT_VEC2(i_xor, V2, V3, V4, T, M2, M3, M4)         \
  S2M_INT_VEC1(i_neg, V2, V3, V4, M2, M3, M4) S2M_INT_VEC1(i_not, V2, V3, V4, M2, M3, M4)
S2M_INT_FAMILY(ivec2, ivec3, ivec4, int, mki2, mki3, mki4)
S2M_INT_FAMILY(uvec2, uvec3, uvec4, unsigned, mku2, mku3, mku4)
/* shifts: the count is always unsigned (scalar or vector of the same width) */
#define S2M_INT_SHIFT(NAME, V2, V3, V4, M2, M3, M4)                                                    \
  S2M_HD V2 NAME(const V2& a, const uvec2& b) { return M2(NAME(a.x, b.x), NAME(a.y, b.y)); }           \
  S2M_HD V3 NAME(const V3& a, const uvec3& b) { return M3(NAME(a.x, b.x), NAME(a.y, b.y), NAME(a.z, b.z)); } \
  S2M_HD V4 NAME(const V4& a, const uvec4& b) { return M4(NAME(a.x, b.x), NAME(a.y, b.y), NAME(a.z, b.z), NAME(a.w, b.w)); } \
  S2M_HD V2 NAME(const V2& a, const ivec2& b) { return M2(NAME(a.x, b.x), NAME(a.y, b.y)); }           \
  S2M_HD V3 NAME(const V3& a, const ivec3& b) { return M3(NAME(a.x, b.x), NAME(a.y, b.y), NAME(a.z, b.z)); } \
  S2M_HD V4 NAME(const V4& a, const ivec4& b) { return M4(NAME(a.x, b.x), NAME(a.y, b.y), NAME(a.z, b.z), NAME(a.w, b.w)); } \
  S2M_HD V2 NAME(const V2& a, unsigned b) { return M2(NAME(a.x, b), NAME(a.y, b)); }                   \
  S2M_HD V3 NAME(const V3& a, unsigned b) { return M3(NAME(a.x, b), NAME(a.y, b), NAME(a.z, b)); }     \
  S2M_HD V4 NAME(const V4& a, unsigned b) { return M4(NAME(a.x, b), NAME(a.y, b), NAME(a.z, b), NAME(a.w, b)); } \
  S2M_HD V2 NAME(const V2& a, int b) { return NAME(a, (unsigned)b); }                                  \
  S2M_HD V3 NAME(const V3& a, int b) { return NAME(a, (unsigned)b); }                                  \
  S2M_HD V4 NAME(const V4& a, int b) { return NAME(a, (unsigned)b); }
S2M_INT_SHIFT(i_shl, ivec2, ivec3, ivec4, mki2, mki3, mki4) S2M_INT_SHIFT(i_shr, ivec2, ivec3, ivec4, mki2, mki3, mki4)
S2M_INT_SHIFT(i_shl, uvec2, uvec3, uvec4, mku2, mku3, mku4) S2M_INT_SHIFT(i_shr, uvec2, uvec3, uvec4, mku2, mku3, mku4)
#undef S2M_INT_SHIFT
/* integer abs / sign / min / max / clamp */
S2M_HD int i_min(int a, int b) { return a < b ? a : b; }
S2M_HD int i_max(int a, int b) { return a > b ? a : b; }
S2M_HD int i_abs(int a) { return a < 0 ? i_neg(a) : a; }
S2M_HD int i_sign(int a) { return a > 0 ? 1 : (a < 0 ? -1 : 0); }
S2M_HD int i_clamp(int x, int lo, int hi) { return i_min(i_max(x, lo), hi); }
S2M_HD unsigned i_min(unsigned a, unsigned b) { return a < b ? a : b; }
S2M_HD unsigned i_max(unsigned a, unsigned b) { return a > b ? a : b; }
S2M_HD unsigned i_abs(unsigned a) { return a; }
S2M_HD unsigned i_clamp(unsigned x, unsigned lo, unsigned hi) { return i_min(i_max(x, lo), hi); }
S2M_INT_VEC2(i_min, ivec2, ivec3, ivec4, int, mki2, mki3, mki4) S2M_INT_VEC2(i_max, ivec2, ivec3, ivec4, int, mki2, mki3, mki4)
S2M_INT_VEC2(i_min, uvec2, uvec3, uvec4, unsigned, mku2, mku3, mku4) S2M_INT_VEC2(i_max, uvec2, uvec3, uvec4, unsigned, mku2, mku3, mku4)
S2M_INT_VEC1(i_abs, ivec2, ivec3, ivec4, mki2, mki3, mki4) S2M_INT_VEC1(i_sign, ivec2, ivec3, ivec4, mki2, mki3, mki4)
S2M_INT_VEC1(i_abs, uvec2, uvec3, uvec4, mku2, mku3, mku4)
/* bit counting (WGSL countOneBits ... firstTrailingBit = GLSL bitCount, bitfieldReverse, findMSB, findLSB): integer
 * operations only, so host and device agree trivially */
S2M_HD unsigned i_countOneBits(unsigned a) {
  a = a - ((a >> 1) & 0x55555555u);
  a = (a & 0x33333333u) + ((a >> 2) & 0x33333333u);
  a = (a + (a >> 4)) & 0x0f0f0f0fu;
  return (a * 0x01010101u) >> 24;
}
S2M_HD unsigned i_reverseBits(unsigned a) {
  a = ((a >> 1) & 0x55555555u) | ((a & 0x55555555u) << 1);
  a = ((a >> 2) & 0x33333333u) | ((a & 0x33333333u) << 2);
  a = ((a >> 4) & 0x0f0f0f0fu) | ((a & 0x0f0f0f0fu) << 4);
  a = ((a >> 8) & 0x00ff00ffu) | ((a & 0x00ff00ffu) << 8);
  return (a >> 16) | (a << 16);
}
S2M_HD unsigned i_countLeadingZeros(unsigned a) {   /* 32 for 0 */
  a |= a >> 1; a |= a >> 2; a |= a >> 4; a |= a >> 8; a |= a >> 16;
  return 32u - i_countOneBits(a);
}
S2M_HD unsigned i_countTrailingZeros(unsigned a) { return i_countOneBits(~a & (a - 1u)); }   /* 32 for 0 */
S2M_HD unsigned i_firstLeadingBit(unsigned a) { return 31u - i_countLeadingZeros(a); }        /* 0xffffffff for 0 */
S2M_HD unsigned i_firstTrailingBit(unsigned a) { return a == 0u ? 0xffffffffu : i_countTrailingZeros(a); }
S2M_HD int i_countOneBits(int a) { return (int)i_countOneBits((unsigned)a); }
S2M_HD int i_reverseBits(int a) { return (int)i_reverseBits((unsigned)a); }
S2M_HD int i_countLeadingZeros(int a) { return (int)i_countLeadingZeros((unsigned)a); }
S2M_HD int i_countTrailingZeros(int a) { return (int)i_countTrailingZeros((unsigned)a); }
S2M_HD int i_firstLeadingBit(int a) { return (int)i_firstLeadingBit((unsigned)(a < 0 ? ~a : a)); }  /* the most significant bit that differs from the sign; -1 for 0 and -1 */
S2M_HD int i_firstTrailingBit(int a) { return (int)i_firstTrailingBit((unsigned)a); }
/* extractBits / insertBits (GLSL bitfieldExtract / bitfieldInsert) with WGSL's clamping: o = min(offset, 32), c = min(count, 32 - o) */
S2M_HD unsigned i_extractBits(unsigned e, unsigned offset, unsigned count) {
  const unsigned o = offset < 32u ? offset : 32u, c = count < 32u - o ? count : 32u - o;
  if (c == 0u) return 0u;
  return (e >> o) & (c == 32u ? 0xffffffffu : ((1u << c) - 1u));   /* o < 32 here: c > 0 */
}
S2M_HD int i_extractBits(int e, unsigned offset, unsigned count) {
  const unsigned o = offset < 32u ? offset : 32u, c = count < 32u - o ? count : 32u - o;
  if (c == 0u) return 0;
  return (int)((unsigned)e << (32u - c - o)) >> (32u - c);        /* sign-extended */
}
S2M_HD unsigned i_insertBits(unsigned e, unsigned newbits, unsigned offset, unsigned count) {
  const unsigned o = offset < 32u ? offset : 32u, c = count < 32u - o ? count : 32u - o;
  if (c == 0u) return e;
  const unsigned mask = (c == 32u ? 0xffffffffu : ((1u << c) - 1u)) << o;
  return (e & ~mask) | ((newbits << o) & mask);
}
S2M_HD int i_insertBits(int e, int newbits, unsigned offset, unsigned count) { return (int)i_insertBits((unsigned)e, (unsigned)newbits, offset, count); }
#define S2M_INT_FIELD(V2, V3, V4, M2, M3, M4)                                                          \
  S2M_HD V2 i_extractBits(const V2& e, unsigned o, unsigned c) { return M2(i_extractBits(e.x, o, c), i_extractBits(e.y, o, c)); } \
  S2M_HD V3 i_extractBits(const V3& e, unsigned o, unsigned c) { return M3(i_extractBits(e.x, o, c), i_extractBits(e.y, o, c), i_extractBits(e.z, o, c)); } \
  S2M_HD V4 i_extractBits(const V4& e, unsigned o, unsigned c) { return M4(i_extractBits(e.x, o, c), i_extractBits(e.y, o, c), i_extractBits(e.z, o, c), i_extractBits(e.w, o, c)); } \
  S2M_HD V2 i_insertBits(const V2& e, const V2& n, unsigned o, unsigned c) { return M2(i_insertBits(e.x, n.x, o, c), i_insertBits(e.y, n.y, o, c)); } \
  S2M_HD V3 i_insertBits(const V3& e, const V3& n, unsigned o, unsigned c) { return M3(i_insertBits(e.x, n.x, o, c), i_insertBits(e.y, n.y, o, c), i_insertBits(e.z, n.z, o, c)); } \
  S2M_HD V4 i_insertBits(const V4& e, const V4& n, unsigned o, unsigned c) { return M4(i_insertBits(e.x, n.x, o, c), i_insertBits(e.y, n.y, o, c), i_insertBits(e.z, n.z, o, c), i_insertBits(e.w, n.w, o, c)); }
S2M_INT_FIELD(ivec2, ivec3, ivec4, mki2, mki3, mki4)
S2M_INT_FIELD(uvec2, uvec3, uvec4, mku2, mku3, mku4)
#undef S2M_INT_FIELD
#define S2M_INT_BITS(NAME)                                                                             \
  S2M_INT_VEC1(NAME, ivec2, ivec3, ivec4, mki2, mki3, mki4) S2M_INT_VEC1(NAME, uvec2, uvec3, uvec4, mku2, mku3, mku4)
S2M_INT_BITS(i_countOneBits) S2M_INT_BITS(i_reverseBits) S2M_INT_BITS(i_countLeadingZeros)
S2M_INT_BITS(i_countTrailingZeros) S2M_INT_BITS(i_firstLeadingBit) S2M_INT_BITS(i_firstTrailingBit)
#undef S2M_INT_BITS
#define S2M_INT_CLAMP(V2, V3, V4, T, M2, M3, M4)                                                       \
  S2M_HD V2 i_clamp(const V2& x, const V2& lo, const V2& hi) { return i_min(i_max(x, lo), hi); }       \
  S2M_HD V3 i_clamp(const V3& x, const V3& lo, const V3& hi) { return i_min(i_max(x, lo), hi); }       \
  S2M_HD V4 i_clamp(const V4& x, const V4& lo, const V4& hi) { return i_min(i_max(x, lo), hi); }       \
  S2M_HD V2 i_clamp(const V2& x, T lo, T hi) { return i_min(i_max(x, lo), hi); }                       \
  S2M_HD V3 i_clamp(const V3& x, T lo, T hi) { return i_min(i_max(x, lo), hi); }                       \
  S2M_HD V4 i_clamp(const V4& x, T lo, T hi) { return i_min(i_max(x, lo), hi); }
S2M_INT_CLAMP(ivec2, ivec3, ivec4, int, mki2, mki3, mki4)
S2M_INT_CLAMP(uvec2, uvec3, uvec4, unsigned, mku2, mku3, mku4)
#undef S2M_INT_CLAMP
#undef S2M_INT_FAMILY
#undef S2M_INT_VEC1
#undef S2M_INT_VEC2

/* ---- conversions */
S2M_HD vec2 to_f(const ivec2& v) { return mk2((float)v.x, (float)v.y); }
S2M_HD vec3 to_f(const ivec3& v) { return mk3((float)v.x, (float)v.y, (float)v.z); }
S2M_HD vec4 to_f(const ivec4& v) { return mk4((float)v.x, (float)v.y, (float)v.z, (float)v.w); }
S2M_HD ivec2 to_i(const vec2& v) { return mki2(s2m_f2int(v.x), s2m_f2int(v.y)); }
S2M_HD ivec3 to_i(const vec3& v) { return mki3(s2m_f2int(v.x), s2m_f2int(v.y), s2m_f2int(v.z)); }
S2M_HD ivec4 to_i(const vec4& v) { return mki4(s2m_f2int(v.x), s2m_f2int(v.y), s2m_f2int(v.z), s2m_f2int(v.w)); }
S2M_HD vec2 to_f(const uvec2& v) { return mk2((float)v.x, (float)v.y); }
S2M_HD vec3 to_f(const uvec3& v) { return mk3((float)v.x, (float)v.y, (float)v.z); }
S2M_HD vec4 to_f(const uvec4& v) { return mk4((float)v.x, (float)v.y, (float)v.z, (float)v.w); }
S2M_HD uvec2 to_u(const vec2& v) { return mku2(s2m_f2uint(v.x), s2m_f2uint(v.y)); }
S2M_HD uvec3 to_u(const vec3& v) { return mku3(s2m_f2uint(v.x), s2m_f2uint(v.y), s2m_f2uint(v.z)); }
S2M_HD uvec4 to_u(const vec4& v) { return mku4(s2m_f2uint(v.x), s2m_f2uint(v.y), s2m_f2uint(v.z), s2m_f2uint(v.w)); }
S2M_HD uvec2 to_u(const ivec2& v) { return mku2((unsigned)v.x, (unsigned)v.y); }
S2M_HD uvec3 to_u(const ivec3& v) { return mku3((unsigned)v.x, (unsigned)v.y, (unsigned)v.z); }
S2M_HD uvec4 to_u(const ivec4& v) { return mku4((unsigned)v.x, (unsigned)v.y, (unsigned)v.z, (unsigned)v.w); }
S2M_HD ivec2 to_i(const uvec2& v) { return mki2((int)v.x, (int)v.y); }
S2M_HD ivec3 to_i(const uvec3& v) { return mki3((int)v.x, (int)v.y, (int)v.z); }
S2M_HD ivec4 to_i(const uvec4& v) { return mki4((int)v.x, (int)v.y, (int)v.z, (int)v.w); }
S2M_HD vec2 to_f(const bvec2& v) { return mk2(v.x ? 1.0f : 0.0f, v.y ? 1.0f : 0.0f); }
S2M_HD vec3 to_f(const bvec3& v) { return mk3(v.x ? 1.0f : 0.0f, v.y ? 1.0f : 0.0f, v.z ? 1.0f : 0.0f); }
S2M_HD vec4 to_f(const bvec4& v) { return mk4(v.x ? 1.0f : 0.0f, v.y ? 1.0f : 0.0f, v.z ? 1.0f : 0.0f, v.w ? 1.0f : 0.0f); }
S2M_HD ivec2 to_i(const bvec2& v) { return mki2(v.x, v.y); }
S2M_HD ivec3 to_i(const bvec3& v) { return mki3(v.x, v.y, v.z); }
S2M_HD ivec4 to_i(const bvec4& v) { return mki4(v.x, v.y, v.z, v.w); }
S2M_HD uvec2 to_u(const bvec2& v) { return mku2(v.x, v.y); }
S2M_HD uvec3 to_u(const bvec3& v) { return mku3(v.x, v.y, v.z); }
S2M_HD uvec4 to_u(const bvec4& v) { return mku4(v.x, v.y, v.z, v.w); }
/* bit reinterpretation (WGSL bitcast<T>, GLSL floatBitsToInt / floatBitsToUint / intBitsToFloat / uintBitsToFloat) */
S2M_HD int bits_i(float a) { return s2m_f2i(a); }
S2M_HD int bits_i(unsigned a) { return (int)a; }
S2M_HD int bits_i(int a) { return a; }
S2M_HD unsigned bits_u(float a) { return (unsigned)s2m_f2i(a); }
S2M_HD unsigned bits_u(int a) { return (unsigned)a; }
S2M_HD unsigned bits_u(unsigned a) { return a; }
S2M_HD float bits_f(int a) { return s2m_i2f(a); }
S2M_HD float bits_f(unsigned a) { return s2m_i2f((int)a); }
S2M_HD float bits_f(float a) { return a; }
#define S2M_BITS_VEC(NAME, R2, R3, R4, M2, M3, M4, A2, A3, A4)                                         \
  S2M_HD R2 NAME(const A2& a) { return M2(NAME(a.x), NAME(a.y)); }                                     \
  S2M_HD R3 NAME(const A3& a) { return M3(NAME(a.x), NAME(a.y), NAME(a.z)); }                          \
  S2M_HD R4 NAME(const A4& a) { return M4(NAME(a.x), NAME(a.y), NAME(a.z), NAME(a.w)); }
S2M_BITS_VEC(bits_i, ivec2, ivec3, ivec4, mki2, mki3, mki4, vec2, vec3, vec4)
S2M_BITS_VEC(bits_i, ivec2, ivec3, ivec4, mki2, mki3, mki4, uvec2, uvec3, uvec4)
S2M_BITS_VEC(bits_u, uvec2, uvec3, uvec4, mku2, mku3, mku4, vec2, vec3, vec4)
S2M_BITS_VEC(bits_u, uvec2, uvec3, uvec4, mku2, mku3, mku4, ivec2, ivec3, ivec4)
S2M_BITS_VEC(bits_f, vec2, vec3, vec4, mk2, mk3, mk4, ivec2, ivec3, ivec4)
S2M_BITS_VEC(bits_f, vec2, vec3, vec4, mk2, mk3, mk4, uvec2, uvec3, uvec4)
#undef S2M_BITS_VEC

#define S2M_SWZ_FAMILY(V2, V3, V4, M2, M3, M4)                                                          \
  S2M_HD V2 swz2(const V2& v, int a, int b) { return M2(cget(v, a), cget(v, b)); }                      \
  S2M_HD V2 swz2(const V3& v, int a, int b) { return M2(cget(v, a), cget(v, b)); }                      \
  S2M_HD V2 swz2(const V4& v, int a, int b) { return M2(cget(v, a), cget(v, b)); }                      \
  S2M_HD V3 swz3(const V2& v, int a, int b, int c) { return M3(cget(v, a), cget(v, b), cget(v, c)); }   \
  S2M_HD V3 swz3(const V3& v, int a, int b, int c) { return M3(cget(v, a), cget(v, b), cget(v, c)); }   \
  S2M_HD V3 swz3(const V4& v, int a, int b, int c) { return M3(cget(v, a), cget(v, b), cget(v, c)); }   \
  S2M_HD V4 swz4(const V2& v, int a, int b, int c, int d) { return M4(cget(v, a), cget(v, b), cget(v, c), cget(v, d)); } \
  S2M_HD V4 swz4(const V3& v, int a, int b, int c, int d) { return M4(cget(v, a), cget(v, b), cget(v, c), cget(v, d)); } \
  S2M_HD V4 swz4(const V4& v, int a, int b, int c, int d) { return M4(cget(v, a), cget(v, b), cget(v, c), cget(v, d)); }
S2M_SWZ_FAMILY(uvec2, uvec3, uvec4, mku2, mku3, mku4)
S2M_SWZ_FAMILY(bvec2, bvec3, bvec4, mkb2, mkb3, mkb4)
#undef S2M_SWZ_FAMILY
S2M_HD ivec3 swz3(const ivec2& v, int a, int b, int c) { return mki3(cget(v, a), cget(v, b), cget(v, c)); }
S2M_HD ivec4 swz4(const ivec2& v, int a, int b, int c, int d) { return mki4(cget(v, a), cget(v, b), cget(v, c), cget(v, d)); }
S2M_HD ivec4 swz4(const ivec3& v, int a, int b, int c, int d) { return mki4(cget(v, a), cget(v, b), cget(v, c), cget(v, d)); }
S2M_HD ivec4 swz4(const ivec4& v, int a, int b, int c, int d) { return mki4(cget(v, a), cget(v, b), cget(v, c), cget(v, d)); }

/* ---- component-wise maps of the scalar builtins in s2m_math.h */
#define S2M_MAP1(NAME, FN)                                                              \
  S2M_HD float NAME(float a) { return FN(a); }                                          \
  S2M_HD vec2 NAME(const vec2& a) { return mk2(FN(a.x), FN(a.y)); }                     \
  S2M_HD vec3 NAME(const vec3& a) { return mk3(FN(a.x), FN(a.y), FN(a.z)); }            \
  S2M_HD vec4 NAME(const vec4& a) { return mk4(FN(a.x), FN(a.y), FN(a.z), FN(a.w)); }
S2M_MAP1(f_abs, s2m_abs)       S2M_MAP1(f_sign, s2m_sign)     S2M_MAP1(f_floor, s2m_floor)
S2M_MAP1(f_ceil, s2m_ceil)     S2M_MAP1(f_trunc, s2m_trunc)   S2M_MAP1(f_round, s2m_round)
S2M_MAP1(f_fract, s2m_fract)   S2M_MAP1(f_sqrt, s2m_sqrt)     S2M_MAP1(f_inversesqrt, s2m_inversesqrt)
/* (sin x, cos x) as a vec2: what optimize.cpp's pair_sin_cos rewrites sin(e) ... cos(e) into */
S2M_HD vec2 f_sincos_pair(float x) { vec2 r; s2m_sincos(x, &r.x, &r.y); return r; }
S2M_MAP1(f_sin, s2m_sin)       S2M_MAP1(f_cos, s2m_cos)       S2M_MAP1(f_tan, s2m_tan)
S2M_MAP1(f_asin, s2m_asin)     S2M_MAP1(f_acos, s2m_acos)     S2M_MAP1(f_atan, s2m_atan)
S2M_MAP1(f_sinh, s2m_sinh)     S2M_MAP1(f_cosh, s2m_cosh)     S2M_MAP1(f_tanh, s2m_tanh)
S2M_MAP1(f_asinh, s2m_asinh)   S2M_MAP1(f_acosh, s2m_acosh)   S2M_MAP1(f_atanh, s2m_atanh)
S2M_MAP1(f_exp, s2m_exp)       S2M_MAP1(f_exp2, s2m_exp2)     S2M_MAP1(f_log, s2m_log)
S2M_MAP1(f_log2, s2m_log2)     S2M_MAP1(f_radians, s2m_radians) S2M_MAP1(f_degrees, s2m_degrees)
#undef S2M_MAP1
S2M_HD float s2m__saturate(float a) { return s2m_clamp(a, 0.0f, 1.0f); }
S2M_HD float f_saturate(float a) { return s2m__saturate(a); }
S2M_HD vec2 f_saturate(const vec2& a) { return mk2(s2m__saturate(a.x), s2m__saturate(a.y)); }
S2M_HD vec3 f_saturate(const vec3& a) { return mk3(s2m__saturate(a.x), s2m__saturate(a.y), s2m__saturate(a.z)); }
S2M_HD vec4 f_saturate(const vec4& a) { return mk4(s2m__saturate(a.x), s2m__saturate(a.y), s2m__saturate(a.z), s2m__saturate(a.w)); }

/* two-operand maps; the scalar-broadcast forms are GLSL's (min(vec,float), pow is vec/vec only) */
#define S2M_MAP2(NAME, FN)                                                                              \
  S2M_HD float NAME(float a, float b) { return FN(a, b); }                                              \
  S2M_HD vec2 NAME(const vec2& a, const vec2& b) { return mk2(FN(a.x, b.x), FN(a.y, b.y)); }            \
  S2M_HD vec3 NAME(const vec3& a, const vec3& b) { return mk3(FN(a.x, b.x), FN(a.y, b.y), FN(a.z, b.z)); } \
  S2M_HD vec4 NAME(const vec4& a, const vec4& b) { return mk4(FN(a.x, b.x), FN(a.y, b.y), FN(a.z, b.z), FN(a.w, b.w)); } \
  S2M_HD vec2 NAME(const vec2& a, float b) { return mk2(FN(a.x, b), FN(a.y, b)); }                      \
  S2M_HD vec3 NAME(const vec3& a, float b) { return mk3(FN(a.x, b), FN(a.y, b), FN(a.z, b)); }          \
  S2M_HD vec4 NAME(const vec4& a, float b) { return mk4(FN(a.x, b), FN(a.y, b), FN(a.z, b), FN(a.w, b)); } \
  S2M_HD vec2 NAME(float a, const vec2& b) { return mk2(FN(a, b.x), FN(a, b.y)); }                      \
  S2M_HD vec3 NAME(float a, const vec3& b) { return mk3(FN(a, b.x), FN(a, b.y), FN(a, b.z)); }          \
  S2M_HD vec4 NAME(float a, const vec4& b) { return mk4(FN(a, b.x), FN(a, b.y), FN(a, b.z), FN(a, b.w)); }
S2M_MAP2(f_min, s2m_min)   S2M_MAP2(f_max, s2m_max)   S2M_MAP2(f_pow, s2m_pow)   S2M_MAP2(f_atan2, s2m_atan2)
S2M_MAP2(f_step, s2m_step) S2M_MAP2(f_mod, s2m_mod_floor) S2M_MAP2(f_rem, s2m_fmod_trunc)
#undef S2M_MAP2

S2M_HD float f_clamp(float x, float lo, float hi) { return s2m_clamp(x, lo, hi); }
S2M_HD vec2 f_clamp(const vec2& x, const vec2& lo, const vec2& hi) { return mk2(s2m_clamp(x.x, lo.x, hi.x), s2m_clamp(x.y, lo.y, hi.y)); }
S2M_HD vec3 f_clamp(const vec3& x, const vec3& lo, const vec3& hi) { return mk3(s2m_clamp(x.x, lo.x, hi.x), s2m_clamp(x.y, lo.y, hi.y), s2m_clamp(x.z, lo.z, hi.z)); }
S2M_HD vec4 f_clamp(const vec4& x, const vec4& lo, const vec4& hi) { return mk4(s2m_clamp(x.x, lo.x, hi.x), s2m_clamp(x.y, lo.y, hi.y), s2m_clamp(x.z, lo.z, hi.z), s2m_clamp(x.w, lo.w, hi.w)); }
S2M_HD vec2 f_clamp(const vec2& x, float lo, float hi) { return mk2(s2m_clamp(x.x, lo, hi), s2m_clamp(x.y, lo, hi)); }
S2M_HD vec3 f_clamp(const vec3& x, float lo, float hi) { return mk3(s2m_clamp(x.x, lo, hi), s2m_clamp(x.y, lo, hi), s2m_clamp(x.z, lo, hi)); }
S2M_HD vec4 f_clamp(const vec4& x, float lo, float hi) { return mk4(s2m_clamp(x.x, lo, hi), s2m_clamp(x.y, lo, hi), s2m_clamp(x.z, lo, hi), s2m_clamp(x.w, lo, hi)); }

S2M_HD float f_mix(float a, float b, float t) { return s2m_mix(a, b, t); }
S2M_HD vec2 f_mix(const vec2& a, const vec2& b, const vec2& t) { return mk2(s2m_mix(a.x, b.x, t.x), s2m_mix(a.y, b.y, t.y)); }
S2M_HD vec3 f_mix(const vec3& a, const vec3& b, const vec3& t) { return mk3(s2m_mix(a.x, b.x, t.x), s2m_mix(a.y, b.y, t.y), s2m_mix(a.z, b.z, t.z)); }
S2M_HD vec4 f_mix(const vec4& a, const vec4& b, const vec4& t) { return mk4(s2m_mix(a.x, b.x, t.x), s2m_mix(a.y, b.y, t.y), s2m_mix(a.z, b.z, t.z), s2m_mix(a.w, b.w, t.w)); }
S2M_HD vec2 f_mix(const vec2& a, const vec2& b, float t) { return mk2(s2m_mix(a.x, b.x, t), s2m_mix(a.y, b.y, t)); }
S2M_HD vec3 f_mix(const vec3& a, const vec3& b, float t) { return mk3(s2m_mix(a.x, b.x, t), s2m_mix(a.y, b.y, t), s2m_mix(a.z, b.z, t)); }
S2M_HD vec4 f_mix(const vec4& a, const vec4& b, float t) { return mk4(s2m_mix(a.x, b.x, t), s2m_mix(a.y, b.y, t), s2m_mix(a.z, b.z, t), s2m_mix(a.w, b.w, t)); }

S2M_HD float f_smoothstep(float lo, float hi, float x) { return s2m_smoothstep(lo, hi, x); }
S2M_HD vec2 f_smoothstep(const vec2& lo, const vec2& hi, const vec2& x) { return mk2(s2m_smoothstep(lo.x, hi.x, x.x), s2m_smoothstep(lo.y, hi.y, x.y)); }
S2M_HD vec3 f_smoothstep(const vec3& lo, const vec3& hi, const vec3& x) { return mk3(s2m_smoothstep(lo.x, hi.x, x.x), s2m_smoothstep(lo.y, hi.y, x.y), s2m_smoothstep(lo.z, hi.z, x.z)); }
S2M_HD vec4 f_smoothstep(const vec4& lo, const vec4& hi, const vec4& x) { return mk4(s2m_smoothstep(lo.x, hi.x, x.x), s2m_smoothstep(lo.y, hi.y, x.y), s2m_smoothstep(lo.z, hi.z, x.z), s2m_smoothstep(lo.w, hi.w, x.w)); }
S2M_HD vec2 f_smoothstep(float lo, float hi, const vec2& x) { return mk2(s2m_smoothstep(lo, hi, x.x), s2m_smoothstep(lo, hi, x.y)); }
S2M_HD vec3 f_smoothstep(float lo, float hi, const vec3& x) { return mk3(s2m_smoothstep(lo, hi, x.x), s2m_smoothstep(lo, hi, x.y), s2m_smoothstep(lo, hi, x.z)); }
S2M_HD vec4 f_smoothstep(float lo, float hi, const vec4& x) { return mk4(s2m_smoothstep(lo, hi, x.x), s2m_smoothstep(lo, hi, x.y), s2m_smoothstep(lo, hi, x.z), s2m_smoothstep(lo, hi, x.w)); }

S2M_HD float f_fma(float a, float b, float c) { return s2m_fma(a, b, c); }
S2M_HD vec2 f_fma(const vec2& a, const vec2& b, const vec2& c) { return mk2(s2m_fma(a.x, b.x, c.x), s2m_fma(a.y, b.y, c.y)); }
S2M_HD vec3 f_fma(const vec3& a, const vec3& b, const vec3& c) { return mk3(s2m_fma(a.x, b.x, c.x), s2m_fma(a.y, b.y, c.y), s2m_fma(a.z, b.z, c.z)); }
S2M_HD vec4 f_fma(const vec4& a, const vec4& b, const vec4& c) { return mk4(s2m_fma(a.x, b.x, c.x), s2m_fma(a.y, b.y, c.y), s2m_fma(a.z, b.z, c.z), s2m_fma(a.w, b.w, c.w)); }

/* ---- geometric */
S2M_HD float f_dot(float a, float b) { return a * b; }
S2M_HD float f_dot(const vec2& a, const vec2& b) { return a.x * b.x + a.y * b.y; }
S2M_HD float f_dot(const vec3& a, const vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
S2M_HD float f_dot(const vec4& a, const vec4& b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
S2M_HD float f_length(float a) { return s2m_abs(a); }
S2M_HD float f_length(const vec2& a) { return s2m_sqrt(f_dot(a, a)); }
S2M_HD float f_length(const vec3& a) { return s2m_sqrt(f_dot(a, a)); }
S2M_HD float f_length(const vec4& a) { return s2m_sqrt(f_dot(a, a)); }
S2M_HD float f_distance(float a, float b) { return s2m_abs(a - b); }
S2M_HD float f_distance(const vec2& a, const vec2& b) { return f_length(a - b); }
S2M_HD float f_distance(const vec3& a, const vec3& b) { return f_length(a - b); }
S2M_HD float f_distance(const vec4& a, const vec4& b) { return f_length(a - b); }
S2M_HD vec2 f_normalize(const vec2& a) { return a / f_length(a); }
S2M_HD vec3 f_normalize(const vec3& a) { return a / f_length(a); }
S2M_HD vec4 f_normalize(const vec4& a) { return a / f_length(a); }
S2M_HD vec3 f_cross(const vec3& a, const vec3& b) {
  return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
S2M_HD vec2 f_reflect(const vec2& i, const vec2& n) { return i - (2.0f * f_dot(n, i)) * n; }
S2M_HD vec3 f_reflect(const vec3& i, const vec3& n) { return i - (2.0f * f_dot(n, i)) * n; }
S2M_HD vec4 f_reflect(const vec4& i, const vec4& n) { return i - (2.0f * f_dot(n, i)) * n; }
/* refract(I, N, eta): k = 1 - eta^2 (1 - dot(N,I)^2); k < 0 ? 0 : eta*I - (eta*dot(N,I) + sqrt(k))*N */
#define S2M_REFRACT(V, ZERO)                                                                           \
  S2M_HD V f_refract(const V& i, const V& n, float eta) {                                              \
    const float d = f_dot(n, i);                                                                       \
    const float k = 1.0f - eta * eta * (1.0f - d * d);                                                 \
    if (k < 0.0f) return ZERO;                                                                         \
    return eta * i - (eta * d + s2m_sqrt(k)) * n;                                                      \
  }                                                                                                    \
  S2M_HD V f_faceforward(const V& n, const V& i, const V& nref) { return f_dot(nref, i) < 0.0f ? n : -n; }
S2M_REFRACT(vec2, splat2(0.0f)) S2M_REFRACT(vec3, splat3(0.0f)) S2M_REFRACT(vec4, splat4(0.0f))
#undef S2M_REFRACT


/* ---- square matrices, column-major (c0 = first column), as in WGSL / GLSL.
 * Pinned: (M*v)_i = sum_j M[j][i]*v[j], summed left to right; v*M = (dot(v,M[0]), dot(v,M[1]), ...);
 * (A*B)[j] = A*B[j]. */
struct mat2 { vec2 c0, c1; };
struct mat3 { vec3 c0, c1, c2; };
struct mat4 { vec4 c0, c1, c2, c3; };
S2M_HD mat2 mkm2(const vec2& a, const vec2& b) { mat2 m; m.c0 = a; m.c1 = b; return m; }
S2M_HD mat3 mkm3(const vec3& a, const vec3& b, const vec3& c) { mat3 m; m.c0 = a; m.c1 = b; m.c2 = c; return m; }
S2M_HD mat4 mkm4(const vec4& a, const vec4& b, const vec4& c, const vec4& d) { mat4 m; m.c0 = a; m.c1 = b; m.c2 = c; m.c3 = d; return m; }
S2M_HD mat2 mkm2(float a, float b, float c, float d) { return mkm2(mk2(a, b), mk2(c, d)); }
S2M_HD mat3 mkm3(float a, float b, float c, float d, float e, float f, float g, float h, float i) { return mkm3(mk3(a, b, c), mk3(d, e, f), mk3(g, h, i)); }
S2M_HD mat4 mkm4(float a, float b, float c, float d, float e, float f, float g, float h, float i, float j, float k, float l, float m, float n, float o, float p) {
  return mkm4(mk4(a, b, c, d), mk4(e, f, g, h), mk4(i, j, k, l), mk4(m, n, o, p));
}
S2M_HD mat2 diagm2(float s) { return mkm2(s, 0.0f, 0.0f, s); }
S2M_HD mat3 diagm3(float s) { return mkm3(s, 0.0f, 0.0f, 0.0f, s, 0.0f, 0.0f, 0.0f, s); }
S2M_HD mat4 diagm4(float s) { return mkm4(s, 0.0f, 0.0f, 0.0f, 0.0f, s, 0.0f, 0.0f, 0.0f, 0.0f, s, 0.0f, 0.0f, 0.0f, 0.0f, s); }
S2M_HD vec2 operator*(const mat2& m, const vec2& v) { return mk2(m.c0.x * v.x + m.c1.x * v.y, m.c0.y * v.x + m.c1.y * v.y); }
S2M_HD vec3 operator*(const mat3& m, const vec3& v) {
  return mk3(m.c0.x * v.x + m.c1.x * v.y + m.c2.x * v.z, m.c0.y * v.x + m.c1.y * v.y + m.c2.y * v.z, m.c0.z * v.x + m.c1.z * v.y + m.c2.z * v.z);
}
S2M_HD vec4 operator*(const mat4& m, const vec4& v) {
  return mk4(m.c0.x * v.x + m.c1.x * v.y + m.c2.x * v.z + m.c3.x * v.w, m.c0.y * v.x + m.c1.y * v.y + m.c2.y * v.z + m.c3.y * v.w,
             m.c0.z * v.x + m.c1.z * v.y + m.c2.z * v.z + m.c3.z * v.w, m.c0.w * v.x + m.c1.w * v.y + m.c2.w * v.z + m.c3.w * v.w);
}
S2M_HD vec2 operator*(const vec2& v, const mat2& m) { return mk2(f_dot(v, m.c0), f_dot(v, m.c1)); }
S2M_HD vec3 operator*(const vec3& v, const mat3& m) { return mk3(f_dot(v, m.c0), f_dot(v, m.c1), f_dot(v, m.c2)); }
S2M_HD vec4 operator*(const vec4& v, const mat4& m) { return mk4(f_dot(v, m.c0), f_dot(v, m.c1), f_dot(v, m.c2), f_dot(v, m.c3)); }
S2M_HD mat2 operator*(const mat2& a, const mat2& b) { return mkm2(a * b.c0, a * b.c1); }
S2M_HD mat3 operator*(const mat3& a, const mat3& b) { return mkm3(a * b.c0, a * b.c1, a * b.c2); }
S2M_HD mat4 operator*(const mat4& a, const mat4& b) { return mkm4(a * b.c0, a * b.c1, a * b.c2, a * b.c3); }
S2M_HD mat2 operator*(const mat2& a, float s) { return mkm2(a.c0 * s, a.c1 * s); }
S2M_HD mat3 operator*(const mat3& a, float s) { return mkm3(a.c0 * s, a.c1 * s, a.c2 * s); }
S2M_HD mat4 operator*(const mat4& a, float s) { return mkm4(a.c0 * s, a.c1 * s, a.c2 * s, a.c3 * s); }
S2M_HD mat2 operator*(float s, const mat2& a) { return mkm2(s * a.c0, s * a.c1); }
S2M_HD mat3 operator*(float s, const mat3& a) { return mkm3(s * a.c0, s * a.c1, s * a.c2); }
S2M_HD mat4 operator*(float s, const mat4& a) { return mkm4(s * a.c0, s * a.c1, s * a.c2, s * a.c3); }
S2M_HD mat2 operator/(const mat2& a, float s) { return mkm2(a.c0 / s, a.c1 / s); }
S2M_HD mat3 operator/(const mat3& a, float s) { return mkm3(a.c0 / s, a.c1 / s, a.c2 / s); }
S2M_HD mat4 operator/(const mat4& a, float s) { return mkm4(a.c0 / s, a.c1 / s, a.c2 / s, a.c3 / s); }
S2M_HD mat2 operator+(const mat2& a, const mat2& b) { return mkm2(a.c0 + b.c0, a.c1 + b.c1); }
S2M_HD mat3 operator+(const mat3& a, const mat3& b) { return mkm3(a.c0 + b.c0, a.c1 + b.c1, a.c2 + b.c2); }
S2M_HD mat4 operator+(const mat4& a, const mat4& b) { return mkm4(a.c0 + b.c0, a.c1 + b.c1, a.c2 + b.c2, a.c3 + b.c3); }
S2M_HD mat2 operator-(const mat2& a, const mat2& b) { return mkm2(a.c0 - b.c0, a.c1 - b.c1); }
S2M_HD mat3 operator-(const mat3& a, const mat3& b) { return mkm3(a.c0 - b.c0, a.c1 - b.c1, a.c2 - b.c2); }
S2M_HD mat4 operator-(const mat4& a, const mat4& b) { return mkm4(a.c0 - b.c0, a.c1 - b.c1, a.c2 - b.c2, a.c3 - b.c3); }
S2M_HD mat2 operator-(const mat2& a) { return mkm2(-a.c0, -a.c1); }
S2M_HD mat3 operator-(const mat3& a) { return mkm3(-a.c0, -a.c1, -a.c2); }
S2M_HD mat4 operator-(const mat4& a) { return mkm4(-a.c0, -a.c1, -a.c2, -a.c3); }
S2M_HD mat2 f_transpose(const mat2& m) { return mkm2(m.c0.x, m.c1.x, m.c0.y, m.c1.y); }
S2M_HD mat3 f_transpose(const mat3& m) { return mkm3(m.c0.x, m.c1.x, m.c2.x, m.c0.y, m.c1.y, m.c2.y, m.c0.z, m.c1.z, m.c2.z); }
S2M_HD mat4 f_transpose(const mat4& m) {
  return mkm4(m.c0.x, m.c1.x, m.c2.x, m.c3.x, m.c0.y, m.c1.y, m.c2.y, m.c3.y, m.c0.z, m.c1.z, m.c2.z, m.c3.z, m.c0.w, m.c1.w, m.c2.w, m.c3.w);
}
S2M_HD float f_determinant(const mat2& m) { return m.c0.x * m.c1.y - m.c1.x * m.c0.y; }
S2M_HD float f_determinant(const mat3& m) { return f_dot(m.c0, f_cross(m.c1, m.c2)); }
/* 4x4: 2x2 minors of the first two and the last two columns (Laplace expansion along column pairs);
 * the same minors give the inverse.  a_ij = row i of column j. */
struct s2m__minors4 { float s0, s1, s2, s3, s4, s5, c0, c1, c2, c3, c4, c5; };
S2M_HD s2m__minors4 s2m__mat4_minors(const mat4& m) {
  s2m__minors4 k;
  k.s0 = m.c0.x * m.c1.y - m.c0.y * m.c1.x; k.s1 = m.c0.x * m.c1.z - m.c0.z * m.c1.x; k.s2 = m.c0.x * m.c1.w - m.c0.w * m.c1.x;
  k.s3 = m.c0.y * m.c1.z - m.c0.z * m.c1.y; k.s4 = m.c0.y * m.c1.w - m.c0.w * m.c1.y; k.s5 = m.c0.z * m.c1.w - m.c0.w * m.c1.z;
  k.c5 = m.c2.z * m.c3.w - m.c2.w * m.c3.z; k.c4 = m.c2.y * m.c3.w - m.c2.w * m.c3.y; k.c3 = m.c2.y * m.c3.z - m.c2.z * m.c3.y;
  k.c2 = m.c2.x * m.c3.w - m.c2.w * m.c3.x; k.c1 = m.c2.x * m.c3.z - m.c2.z * m.c3.x; k.c0 = m.c2.x * m.c3.y - m.c2.y * m.c3.x;
  return k;
}
S2M_HD float f_determinant(const mat4& m) {
  const s2m__minors4 k = s2m__mat4_minors(m);
  return k.s0 * k.c5 - k.s1 * k.c4 + k.s2 * k.c3 + k.s3 * k.c2 - k.s4 * k.c1 + k.s5 * k.c0;
}
/* inverse = adjugate * (1 / determinant); a singular matrix gives inf / NaN entries (GLSL: undefined) */
S2M_HD mat2 f_inverse(const mat2& m) {
  const float r = 1.0f / f_determinant(m);
  return mkm2(m.c1.y * r, -m.c0.y * r, -m.c1.x * r, m.c0.x * r);
}
S2M_HD mat3 f_inverse(const mat3& m) {
  const vec3 r0 = f_cross(m.c1, m.c2), r1 = f_cross(m.c2, m.c0), r2 = f_cross(m.c0, m.c1);  /* rows of the adjugate */
  const float r = 1.0f / f_dot(m.c0, r0);
  return mkm3(r0.x * r, r1.x * r, r2.x * r, r0.y * r, r1.y * r, r2.y * r, r0.z * r, r1.z * r, r2.z * r);
}
S2M_HD mat4 f_inverse(const mat4& m) {
  const s2m__minors4 k = s2m__mat4_minors(m);
  const float r = 1.0f / (k.s0 * k.c5 - k.s1 * k.c4 + k.s2 * k.c3 + k.s3 * k.c2 - k.s4 * k.c1 + k.s5 * k.c0);
  /* inv[i][j] (row i, column j); columns of the result are listed one after the other */
  return mkm4(
      ( m.c1.y * k.c5 - m.c1.z * k.c4 + m.c1.w * k.c3) * r, (-m.c0.y * k.c5 + m.c0.z * k.c4 - m.c0.w * k.c3) * r,
      ( m.c3.y * k.s5 - m.c3.z * k.s4 + m.c3.w * k.s3) * r, (-m.c2.y * k.s5 + m.c2.z * k.s4 - m.c2.w * k.s3) * r,
      (-m.c1.x * k.c5 + m.c1.z * k.c2 - m.c1.w * k.c1) * r, ( m.c0.x * k.c5 - m.c0.z * k.c2 + m.c0.w * k.c1) * r,
      (-m.c3.x * k.s5 + m.c3.z * k.s2 - m.c3.w * k.s1) * r, ( m.c2.x * k.s5 - m.c2.z * k.s2 + m.c2.w * k.s1) * r,
      ( m.c1.x * k.c4 - m.c1.y * k.c2 + m.c1.w * k.c0) * r, (-m.c0.x * k.c4 + m.c0.y * k.c2 - m.c0.w * k.c0) * r,
      ( m.c3.x * k.s4 - m.c3.y * k.s2 + m.c3.w * k.s0) * r, (-m.c2.x * k.s4 + m.c2.y * k.s2 - m.c2.w * k.s0) * r,
      (-m.c1.x * k.c3 + m.c1.y * k.c1 - m.c1.z * k.c0) * r, ( m.c0.x * k.c3 - m.c0.y * k.c1 + m.c0.z * k.c0) * r,
      (-m.c3.x * k.s3 + m.c3.y * k.s1 - m.c3.z * k.s0) * r, ( m.c2.x * k.s3 - m.c2.y * k.s1 + m.c2.z * k.s0) * r);
}

/* ---- comparisons / selection */
#define S2M_CMP(NAME, OP)                                                                                \
  S2M_HD bvec2 NAME(const vec2& a, const vec2& b) { return mkb2(a.x OP b.x, a.y OP b.y); }               \
  S2M_HD bvec3 NAME(const vec3& a, const vec3& b) { return mkb3(a.x OP b.x, a.y OP b.y, a.z OP b.z); }   \
  S2M_HD bvec4 NAME(const vec4& a, const vec4& b) { return mkb4(a.x OP b.x, a.y OP b.y, a.z OP b.z, a.w OP b.w); }
S2M_CMP(v_lt, <) S2M_CMP(v_le, <=) S2M_CMP(v_gt, >) S2M_CMP(v_ge, >=) S2M_CMP(v_eq, ==) S2M_CMP(v_ne, !=)
#undef S2M_CMP
#define S2M_CMP_FAMILY(NAME, OP, V2, V3, V4)                                                             \
  S2M_HD bvec2 NAME(const V2& a, const V2& b) { return mkb2(a.x OP b.x, a.y OP b.y); }                   \
  S2M_HD bvec3 NAME(const V3& a, const V3& b) { return mkb3(a.x OP b.x, a.y OP b.y, a.z OP b.z); }       \
  S2M_HD bvec4 NAME(const V4& a, const V4& b) { return mkb4(a.x OP b.x, a.y OP b.y, a.z OP b.z, a.w OP b.w); }
S2M_CMP_FAMILY(v_lt, <, ivec2, ivec3, ivec4) S2M_CMP_FAMILY(v_le, <=, ivec2, ivec3, ivec4) S2M_CMP_FAMILY(v_gt, >, ivec2, ivec3, ivec4)
S2M_CMP_FAMILY(v_ge, >=, ivec2, ivec3, ivec4) S2M_CMP_FAMILY(v_eq, ==, ivec2, ivec3, ivec4) S2M_CMP_FAMILY(v_ne, !=, ivec2, ivec3, ivec4)
S2M_CMP_FAMILY(v_lt, <, uvec2, uvec3, uvec4) S2M_CMP_FAMILY(v_le, <=, uvec2, uvec3, uvec4) S2M_CMP_FAMILY(v_gt, >, uvec2, uvec3, uvec4)
S2M_CMP_FAMILY(v_ge, >=, uvec2, uvec3, uvec4) S2M_CMP_FAMILY(v_eq, ==, uvec2, uvec3, uvec4) S2M_CMP_FAMILY(v_ne, !=, uvec2, uvec3, uvec4)
S2M_CMP_FAMILY(v_eq, ==, bvec2, bvec3, bvec4) S2M_CMP_FAMILY(v_ne, !=, bvec2, bvec3, bvec4)
S2M_CMP_FAMILY(v_and, &&, bvec2, bvec3, bvec4) S2M_CMP_FAMILY(v_or, ||, bvec2, bvec3, bvec4)
#undef S2M_CMP_FAMILY
S2M_HD bvec2 v_not(const bvec2& a) { return mkb2(!a.x, !a.y); }
S2M_HD bvec3 v_not(const bvec3& a) { return mkb3(!a.x, !a.y, !a.z); }
S2M_HD bvec4 v_not(const bvec4& a) { return mkb4(!a.x, !a.y, !a.z, !a.w); }
S2M_HD bool b_all(bool a) { return a; }
S2M_HD bool b_all(const bvec2& a) { return a.x && a.y; }
S2M_HD bool b_all(const bvec3& a) { return a.x && a.y && a.z; }
S2M_HD bool b_all(const bvec4& a) { return a.x && a.y && a.z && a.w; }
S2M_HD bool b_any(bool a) { return a; }
S2M_HD bool b_any(const bvec2& a) { return a.x || a.y; }
S2M_HD bool b_any(const bvec3& a) { return a.x || a.y || a.z; }
S2M_HD bool b_any(const bvec4& a) { return a.x || a.y || a.z || a.w; }
/* WGSL select(f, t, cond) */
/* any other type (structs, arrays, matrices) with a scalar condition: GLSL's ?: on aggregates arrives here */
template <class T> S2M_HD T f_select(const T& f, const T& t, bool c) { return c ? t : f; }
S2M_HD float f_select(float f, float t, bool c) { return c ? t : f; }
S2M_HD int f_select(int f, int t, bool c) { return c ? t : f; }
S2M_HD unsigned f_select(unsigned f, unsigned t, bool c) { return c ? t : f; }
S2M_HD bool f_select(bool f, bool t, bool c) { return c ? t : f; }
S2M_HD vec2 f_select(const vec2& f, const vec2& t, bool c) { return c ? t : f; }
S2M_HD vec3 f_select(const vec3& f, const vec3& t, bool c) { return c ? t : f; }
S2M_HD vec4 f_select(const vec4& f, const vec4& t, bool c) { return c ? t : f; }
S2M_HD vec2 f_select(const vec2& f, const vec2& t, const bvec2& c) { return mk2(c.x ? t.x : f.x, c.y ? t.y : f.y); }
S2M_HD vec3 f_select(const vec3& f, const vec3& t, const bvec3& c) { return mk3(c.x ? t.x : f.x, c.y ? t.y : f.y, c.z ? t.z : f.z); }
S2M_HD vec4 f_select(const vec4& f, const vec4& t, const bvec4& c) { return mk4(c.x ? t.x : f.x, c.y ? t.y : f.y, c.z ? t.z : f.z, c.w ? t.w : f.w); }
#define S2M_SELECT_FAMILY(V2, V3, V4, M2, M3, M4)                                                        \
  S2M_HD V2 f_select(const V2& f, const V2& t, bool c) { return c ? t : f; }                             \
  S2M_HD V3 f_select(const V3& f, const V3& t, bool c) { return c ? t : f; }                             \
  S2M_HD V4 f_select(const V4& f, const V4& t, bool c) { return c ? t : f; }                             \
  S2M_HD V2 f_select(const V2& f, const V2& t, const bvec2& c) { return M2(c.x ? t.x : f.x, c.y ? t.y : f.y); } \
  S2M_HD V3 f_select(const V3& f, const V3& t, const bvec3& c) { return M3(c.x ? t.x : f.x, c.y ? t.y : f.y, c.z ? t.z : f.z); } \
  S2M_HD V4 f_select(const V4& f, const V4& t, const bvec4& c) { return M4(c.x ? t.x : f.x, c.y ? t.y : f.y, c.z ? t.z : f.z, c.w ? t.w : f.w); }
S2M_SELECT_FAMILY(ivec2, ivec3, ivec4, mki2, mki3, mki4)
S2M_SELECT_FAMILY(uvec2, uvec3, uvec4, mku2, mku3, mku4)
S2M_SELECT_FAMILY(bvec2, bvec3, bvec4, mkb2, mkb3, mkb4)
#undef S2M_SELECT_FAMILY


/* ---------------------------------------------------------------- arrays and dynamic indexing
 * array<T, N> / T[N] of the shading languages.  An out-of-range index reads / writes the nearest
 * valid element (the `Restrict` bounds-check policy wgpu applies by default). */
template <class T, int N> struct s2m_array { T v[N]; };
S2M_HD int s2m_clamp_index(int i, int n) { return i < 0 ? 0 : (i >= n ? n - 1 : i); }
S2M_HD int s2m_clamp_index(unsigned i, int n) { return i >= (unsigned)n ? n - 1 : (int)i; }
template <class T, int N, class I> S2M_HD T& s2m_at(s2m_array<T, N>& a, I i) { return a.v[s2m_clamp_index(i, N)]; }
template <class T, int N, class I> S2M_HD const T& s2m_at(const s2m_array<T, N>& a, I i) { return a.v[s2m_clamp_index(i, N)]; }
#define S2M_VAT(V, T, N) \
  template <class I> S2M_HD T& s2m_at(V& v, I i) { return (&v.x)[s2m_clamp_index(i, N)]; } \
  template <class I> S2M_HD const T& s2m_at(const V& v, I i) { return (&v.x)[s2m_clamp_index(i, N)]; }
S2M_VAT(vec2, float, 2) S2M_VAT(vec3, float, 3) S2M_VAT(vec4, float, 4)
S2M_VAT(ivec2, int, 2) S2M_VAT(ivec3, int, 3) S2M_VAT(ivec4, int, 4)
S2M_VAT(uvec2, unsigned, 2) S2M_VAT(uvec3, unsigned, 3) S2M_VAT(uvec4, unsigned, 4)
S2M_VAT(bvec2, bool, 2) S2M_VAT(bvec3, bool, 3) S2M_VAT(bvec4, bool, 4)
#undef S2M_VAT
#define S2M_MAT_AT(M, V, N) \
  template <class I> S2M_HD V& s2m_at(M& m, I i) { return (&m.c0)[s2m_clamp_index(i, N)]; } \
  template <class I> S2M_HD const V& s2m_at(const M& m, I i) { return (&m.c0)[s2m_clamp_index(i, N)]; }
S2M_MAT_AT(mat2, vec2, 2) S2M_MAT_AT(mat3, vec3, 3) S2M_MAT_AT(mat4, vec4, 4)
#undef S2M_MAT_AT

}  // namespace s2m
#endif /* S2M_VEC_H_ */
