// oracle/sdf_examples.h -- TEST INFRASTRUCTURE ONLY (see oracle.cpp header).
//
// Hand transcription of the reference's SDF libraries and example inputs into plain scalar
// C++ (strict f32, no FMA contraction, operation order exactly as written in the WGSL/GLSL).
// Independent of the product's front-end/emitter on purpose: the parity tests compare what the
// emitter + NVRTC produce from the *text* of these files against this transcription.
//
// Float semantics the reference leaves to the driver compiler, pinned here:
//   dot(a,b)      = a.x*b.x + a.y*b.y (+ a.z*b.z), summed left to right, no FMA
//   length(v)     = sqrt(dot(v,v));  normalize(v) = v / length(v);  distance(a,b) = length(a-b)
//   clamp(x,l,h)  = min(max(x,l),h)            (WGSL spec)
//   mix(a,b,t)    = a*(1-t) + b*t              (WGSL spec)
//   min/max       = IEEE minimumNumber/maximumNumber (NaN operand ignored, -0 < +0)
//   module-scope `const X = <abstract float expr>` is folded in f64 and rounded to f32 on use
//                   (naga's abstract-float constant evaluation)
//   sin cos atan asin pow log = sdf2mesh_b200/csrc/s2m_math.h (the engine's pinned math)
#pragma once
#include "../sdf2mesh_b200/csrc/s2m_math.h"

namespace osdf {

struct V2 { float x, y; };
struct V3 { float x, y, z; };

static inline float omin(float a, float b) {
  if (a != a) return b;
  if (b != b) return a;
  if (a == b) return (s2m_f2i(a) < 0) ? a : b;  // -0 before +0
  return a < b ? a : b;
}
static inline float omax(float a, float b) {
  if (a != a) return b;
  if (b != b) return a;
  if (a == b) return (s2m_f2i(a) < 0) ? b : a;  // +0 before -0
  return a > b ? a : b;
}
static inline float oabs(float a) { return a < 0.0f ? -a : (a == 0.0f ? 0.0f : a); }
static inline float oclamp(float x, float lo, float hi) { return omin(omax(x, lo), hi); }
static inline float omix(float a, float b, float t) { return a * (1.0f - t) + b * t; }
static inline float dot2(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
static inline float dot3(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline float len2(V2 a) { return sqrtf(dot2(a, a)); }
static inline float len3(V3 a) { return sqrtf(dot3(a, a)); }
static inline V3 sub3(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 add3(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3 scale3(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }

// ---- /root/reference/src/sdf3d_primitives.wgsl
// :7-11  sdf3d_box
static inline float sdf3d_box(V3 p, V3 b) {
  V3 q = {oabs(p.x) - 0.5f * b.x, oabs(p.y) - 0.5f * b.y, oabs(p.z) - 0.5f * b.z};
  V3 m = {omax(q.x, 0.0f), omax(q.y, 0.0f), omax(q.z, 0.0f)};
  return len3(m) + omin(omax(q.x, omax(q.y, q.z)), 0.0f);
}
// :13-17 sdf3d_cylinder
static inline float sdf3d_cylinder(V3 p, float h, float r) {
  V2 d = {oabs(len2(V2{p.x, p.z})) - r, oabs(p.y) - h};
  return omin(omax(d.x, d.y), 0.0f) + len2(V2{omax(d.x, 0.0f), omax(d.y, 0.0f)});
}
// :19-25 sdf3d_capsule
static inline float sdf3d_capsule(V3 p, V3 a, V3 b, float r) {
  V3 pa = sub3(p, a);
  V3 ba = sub3(b, a);
  float h = oclamp(dot3(pa, ba) / dot3(ba, ba), 0.0f, 1.0f);
  return len3(sub3(pa, scale3(ba, h))) - r;
}
// :27-30 sdf3d_sphere
static inline float sdf3d_sphere(V3 p, float s) { return len3(p) - s; }
// :32-36 sdf3d_torus
static inline float sdf3d_torus(V3 p, V2 t) {
  V2 q = {len2(V2{p.x, p.z}) - t.x, p.y};
  return len2(q) - t.y;
}

// ---- /root/reference/src/sdf_op.wgsl
// :7-11
static inline float sdf_op_smooth_union(float d1, float d2, float k) {
  float h = oclamp(0.5f + 0.5f * (d2 - d1) / k, 0.0f, 1.0f);
  return omix(d2, d1, h) - k * h * (1.0f - h);
}
// :13-17
static inline float sdf_op_smooth_intersection(float d1, float d2, float k) {
  float h = oclamp(0.5f - 0.5f * (d2 - d1) / k, 0.0f, 1.0f);
  return omix(d2, d1, h) + k * h * (1.0f - h);
}
// :19-23
static inline float sdf_op_smooth_subtraction(float d1, float d2, float k) {
  float h = oclamp(0.5f - 0.5f * (d2 + d1) / k, 0.0f, 1.0f);
  return omix(d2, -d1, h) + k * h * (1.0f - h);
}

// ---- /root/reference/examples/torus.sdf3d:3-5
static inline float sdf_torus(V3 p) { return sdf3d_torus(p, V2{0.5f, 0.2f}); }

// ---- /root/reference/examples/martin_cube.sdf3d
namespace martin {
static const double CUBE_WIDTH = 1.0, LETTER_THICKNESS = 0.12, SMOOTHNESS = 0.02, C = CUBE_WIDTH; // :4-8
#define AF(x) ((float)(x)) /* abstract-float constant expression -> f32 */
// :10-28 (shared shape with p_key's letter_p_arc; K is the module constant)
static inline float arc(V3 p, float ra, float rb, double K) {
  V3 pp = {p.y, p.x, p.z};
  pp.x = oabs(pp.x - AF(K * 0.2));
  pp.y += AF(K * 0.25);
  V3 h = {0.0f, AF(K * 0.15), 0.0f};
  V3 nh = {AF(-0.0), AF(-(K * 0.15)), AF(-0.0)};
  V3 q = {pp.x - oclamp(pp.x, nh.x, h.x), pp.y - oclamp(pp.y, nh.y, h.y), pp.z - oclamp(pp.z, nh.z, h.z)};
  V2 sc = {1.0f, 0.0f};
  float k = 0.0f;
  if (0.0f > sc.x * q.y) k = dot2(V2{q.x, q.y}, sc);
  else k = len2(V2{q.x, q.y});
  return sqrtf(dot3(q, q) + ra * ra - 2.0f * ra * k) - rb;
}
static inline float segment(V3 p, V3 a, V3 b) { return sdf3d_capsule(p, a, b, AF(LETTER_THICKNESS)); } // :31-33
static inline float letter_m(V3 p) { // :36-49
  V3 q = {p.y, p.x, p.z};
  q.y *= -1.0f;
  return omin(
      omin(segment(q, V3{AF(-C * 0.45), AF(-C * 0.5), AF(C)}, V3{AF(-C * 0.45), AF(C * 0.5), AF(C)}),
           segment(q, V3{AF(-C * 0.45), AF(-C * 0.5), AF(C)}, V3{0.0f, 0.0f, AF(C)})),
      omin(segment(q, V3{0.0f, 0.0f, AF(C)}, V3{AF(C * 0.45), AF(-C * 0.5), AF(C)}),
           segment(q, V3{AF(C * 0.45), AF(-C * 0.5), AF(C)}, V3{AF(C * 0.45), AF(C * 0.5), AF(C)})));
}
static inline float letter_a(V3 p) { // :52-63
  V3 q = {p.z, p.y, p.x};
  q = V3{-q.y, -q.x, q.z};
  return omin(omin(segment(q, V3{0.0f, AF(-C * 0.5), AF(-C)}, V3{AF(-C * 0.4), AF(C * 0.5), AF(-C)}),
                   segment(q, V3{0.0f, AF(-C * 0.5), AF(-C)}, V3{AF(C * 0.4), AF(C * 0.5), AF(-C)})),
              segment(q, V3{AF(-C * 0.2), AF(C * 0.1), AF(-C)}, V3{AF(C * 0.2), AF(C * 0.1), AF(-C)}));
}
static inline float letter_r(V3 p) { // :66-77
  V3 q = p;
  q.y *= -1.0f;
  q.x -= AF(C * 0.15);
  V3 a = sub3(V3{q.x, q.z, q.y}, V3{0.0f, 0.0f, AF(C)});
  return omin(arc(a, AF(C * 0.3), AF(LETTER_THICKNESS), C),
              omin(segment(q, V3{AF(-C * 0.4), AF(C), AF(C * 0.5)}, V3{AF(-C * 0.4), AF(C), AF(-C * 0.5)}),
                   segment(q, V3{AF(-C * 0.15), AF(C), AF(-C * 0.1)}, V3{AF(C * 0.15), AF(C), AF(-C * 0.5)})));
}
static inline float letter_t(V3 p) { // :81-86
  return omin(segment(p, V3{AF(C), 0.0f, AF(C * 0.5)}, V3{AF(C), 0.0f, AF(-C * 0.5)}),
              segment(p, V3{AF(C), AF(-C * 0.4), AF(C * 0.5)}, V3{AF(C), AF(C * 0.4), AF(C * 0.5)}));
}
static inline float letter_i(V3 p) { // :90-94
  V3 q = {p.y, p.x, p.z};
  q.x *= -1.0f;
  return segment(q, V3{AF(-C), 0.0f, AF(C * 0.5)}, V3{AF(-C), 0.0f, AF(-C * 0.5)});
}
static inline float letter_n(V3 p) { // :98-109
  V3 q = {p.x, p.z, p.y};
  return omin(segment(q, V3{AF(-C * 0.4), AF(-C), AF(C * 0.5)}, V3{AF(-C * 0.4), AF(-C), AF(-C * 0.5)}),
              omin(segment(q, V3{AF(C * 0.4), AF(-C), AF(C * 0.5)}, V3{AF(C * 0.4), AF(-C), AF(-C * 0.5)}),
                   segment(q, V3{AF(-C * 0.4), AF(-C), AF(-C * 0.5)}, V3{AF(C * 0.4), AF(-C), AF(C * 0.5)})));
}
static inline float sdf(V3 p) { // :111-122
  float cube = sdf_op_smooth_intersection(
      sdf3d_box(p, V3{AF(CUBE_WIDTH), AF(CUBE_WIDTH), AF(CUBE_WIDTH)}),
      sdf3d_sphere(p, AF(CUBE_WIDTH * 1.40)), AF(SMOOTHNESS));
  float letter = omin(omin(letter_m(p), letter_a(p)),
                      omin(omin(letter_r(p), letter_t(p)), omin(letter_i(p), letter_n(p))));
  return sdf_op_smooth_subtraction(letter, cube, AF(SMOOTHNESS));
}
}  // namespace martin

// ---- /root/reference/examples/p_key.sdf3d
namespace pkey {
static const double KEY_SIZE = 15.0, KEY_ELEVATION = 2.0, PLATE_HEIGHT = 1.5, LETTER_THICKNESS = 1.0; // :4-7
static inline float segment(V3 p, V3 a, V3 b) { return sdf3d_capsule(p, a, b, AF(LETTER_THICKNESS)); } // :29-31
static inline float letter_p(V3 p) { // :33-41
  V3 q = p;
  q.y *= -1.0f;
  q.x -= AF(KEY_SIZE * 0.15);
  return omin(martin::arc(V3{q.x, q.z, q.y}, AF(KEY_SIZE * 0.3), AF(LETTER_THICKNESS), KEY_SIZE), // :9-27
              segment(q, V3{AF(-KEY_SIZE * 0.4), 0.0f, AF(KEY_SIZE * 0.5)},
                      V3{AF(-KEY_SIZE * 0.4), 0.0f, AF(-KEY_SIZE * 0.5)}));
}
static inline float sdf(V3 p) { // :43-52
  V3 a = sub3(scale3(V3{p.x, p.z, p.y}, 2.5f), V3{0.0f, AF(KEY_ELEVATION * 2.5 + LETTER_THICKNESS), 0.0f});
  V3 c = sub3(V3{p.y, p.z, p.x}, V3{0.0f, AF((KEY_ELEVATION + PLATE_HEIGHT) * 0.5), 0.0f});
  return sdf_op_smooth_subtraction(
      letter_p(a),
      sdf_op_smooth_union(sdf3d_box(p, V3{AF(KEY_SIZE), AF(KEY_SIZE), AF(PLATE_HEIGHT)}),
                          sdf3d_cylinder(c, AF(KEY_ELEVATION - PLATE_HEIGHT), AF(KEY_SIZE * 0.35)),
                          AF(KEY_SIZE * 0.15)),
      0.0f);
}
}  // namespace pkey

// ---- /root/reference/examples/mandelmesh.frag:3-28 (GLSL: every literal is f32)
static inline float sdf_mandelbulb(V3 pin) {
  V3 p = {pin.x, pin.z, pin.y};  // p.xyz = p.xzy
  V3 z = p;
  float power = 8.0f;
  float r = 0.0f, theta = 0.0f, phi = 0.0f;  // naga zero-initialises locals
  float dr = 1.0f;
  for (int i = 0; i < 5; ++i) {
    r = len3(z);
    if (r > 2.0f) continue;
    theta = s2m_atan(z.y / z.x);
    phi = s2m_asin(z.z / r);
    dr = s2m_pow(r, power - 1.0f) * dr * power + 1.0f;
    r = s2m_pow(r, power);
    theta = theta * power;
    phi = phi * power;
    V3 d = {s2m_cos(theta) * s2m_cos(phi), s2m_sin(theta) * s2m_cos(phi), s2m_sin(phi)};
    z = add3(scale3(d, r), p);
  }
  return 0.5f * s2m_log(r) * r / dr - 0.003f;
}

// ---- /root/reference/src/shadertoy.rs:411-442 (the test_naga GLSL shader; iTime uniform = 0)
static inline float sdf_naga_sphere(V3 p) {
  V3 c = {0.0f, 0.0f, 0.0f};
  const float r = 1.0f;
  float sphere_0 = len3(sub3(p, c)) - r;  // distance(p, c) - r
  float iTime = 0.0f;
  float displacement =
      s2m_sin(5.0f * p.x) * s2m_sin(5.0f * p.y) * s2m_sin(5.0f * p.z) * 0.25f * s2m_sin(2.0f * iTime);
  return sphere_0 + displacement;
}

enum SdfId { SDF_TORUS = 0, SDF_MARTIN_CUBE = 1, SDF_P_KEY = 2, SDF_MANDELBULB = 3, SDF_NAGA_SPHERE = 4, SDF_COUNT, SDF_PLUGIN = 99 };

// SDF_PLUGIN: a function supplied by the test (oracle_set_plugin) -- the front-end's emitted C++ compiled
// for the host by tests/support/host_eval.py -- so that the oracle's cell loop, scan and quad assembly
// can be compared with the GPU path for shaders that have no hand transcription here.
typedef float (*PluginFn)(float, float, float);
static PluginFn g_plugin = nullptr;

static inline float eval(int id, V3 p) {
  switch (id) {
    case SDF_TORUS: return sdf_torus(p);
    case SDF_MARTIN_CUBE: return martin::sdf(p);
    case SDF_P_KEY: return pkey::sdf(p);
    case SDF_MANDELBULB: return sdf_mandelbulb(p);
    case SDF_NAGA_SPHERE: return sdf_naga_sphere(p);
    case SDF_PLUGIN: return g_plugin ? g_plugin(p.x, p.y, p.z) : 0.0f;
  }
  return 0.0f;
}

}  // namespace osdf
