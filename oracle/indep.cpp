// oracle/indep.cpp -- INDEPENDENT EVALUATION FOR THE TOLERANCE-MODE PARITY CHECK.
// TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rules as oracle.cpp: only tests/, smoke() and
// bench.py's checker legs may build, load or call it).
//
// Why it exists.  oracle.cpp shares csrc/s2m_math.h with the device code, so "bit-exact vs the
// oracle" says nothing about the transcendental functions themselves.  This file evaluates the
// example SDFs WITHOUT any s2m_* code: a second transcription, templated on the scalar type, run as
//   T = double  with the C library's f64 sin/cos/atan/asin/pow/log/sqrt  ("ground truth": what the
//               shader's real-valued function is, to ~1e-16), and as
//   T = float   with the C library's f32 sinf/cosf/atanf/asinf/powf/logf/sqrtf ("another driver":
//               correctly-ordered f32 arithmetic with somebody else's transcendental functions --
//               the situation of the reference's wgpu path, whose functions come from the
//               Vulkan driver's shader compiler).
// and then applies the reference's cell rule in T (crossings, count >= 2, mean-of-crossings
// position; dualcontour.wgsl:22-43, :57-69, :72-131).  The corner COORDINATES are always the
// reference's f32 ones (cell min = bmin + size*f32(i), max = min + size; :22-27): those are plain
// IEEE multiplications and additions and do not depend on a driver.
//
// BASELINE.json:north_star states the acceptance rule this supports: "identical set of active
// cells and identical quad connectivity (bit-exact, excluding and listing corners with
// |sdf| < 1e-6 voxel), and vertex positions within 1e-4 voxel".  tests/support/tolerance.py does
// the set logic on what indep_run returns.
//
// Transcribed from /root/reference/examples/{torus,martin_cube,p_key}.sdf3d, mandelmesh.frag,
// /root/reference/src/sdf3d_primitives.wgsl, sdf_op.wgsl and shadertoy.rs:411-442 -- not from
// oracle/sdf_examples.h.  Module-scope WGSL constants are abstract floats: expressions over them
// are folded in f64 and rounded to f32 where they meet an f32 (naga), in BOTH instantiations, so
// that T = double evaluates the same real function the f32 shader approximates.
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

namespace {

template <class T> struct M;  // the math library of scalar type T
template <> struct M<double> {
  static double sqrt(double x) { return std::sqrt(x); }
  static double sin(double x) { return std::sin(x); }
  static double cos(double x) { return std::cos(x); }
  static double atan(double x) { return std::atan(x); }
  static double asin(double x) { return std::asin(x); }
  static double pow(double x, double y) { return std::pow(x, y); }
  static double log(double x) { return std::log(x); }
  static double fmin(double a, double b) { return std::fmin(a, b); }
  static double fmax(double a, double b) { return std::fmax(a, b); }
  static double abs(double a) { return std::fabs(a); }
};
template <> struct M<float> {
  static float sqrt(float x) { return sqrtf(x); }
  static float sin(float x) { return sinf(x); }
  static float cos(float x) { return cosf(x); }
  static float atan(float x) { return atanf(x); }
  static float asin(float x) { return asinf(x); }
  static float pow(float x, float y) { return powf(x, y); }
  static float log(float x) { return logf(x); }
  static float fmin(float a, float b) { return fminf(a, b); }
  static float fmax(float a, float b) { return fmaxf(a, b); }
  static float abs(float a) { return fabsf(a); }
};

// an abstract-float constant expression meeting an f32: round once to f32, then widen to T
template <class T> inline T K(double abstract_value) { return (T)(float)abstract_value; }

template <class T> struct P2 { T x, y; };
template <class T> struct P3 { T x, y, z; };
template <class T> inline P3<T> operator-(P3<T> a, P3<T> b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <class T> inline P3<T> operator+(P3<T> a, P3<T> b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <class T> inline P3<T> operator*(P3<T> a, T s) { return {a.x * s, a.y * s, a.z * s}; }
template <class T> inline T dot(P3<T> a, P3<T> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <class T> inline T dot(P2<T> a, P2<T> b) { return a.x * b.x + a.y * b.y; }
template <class T> inline T length(P3<T> a) { return M<T>::sqrt(dot(a, a)); }
template <class T> inline T length(P2<T> a) { return M<T>::sqrt(dot(a, a)); }
template <class T> inline T clamp(T x, T lo, T hi) { return M<T>::fmin(M<T>::fmax(x, lo), hi); }  // WGSL: min(max(e, low), high)
template <class T> inline T mix(T a, T b, T t) { return a * ((T)1 - t) + b * t; }                 // WGSL: e1*(1-e3) + e2*e3

// ---- src/sdf3d_primitives.wgsl
template <class T> T prim_box(P3<T> p, P3<T> b) {  // :7-11
  P3<T> q = {M<T>::abs(p.x) - (T)0.5 * b.x, M<T>::abs(p.y) - (T)0.5 * b.y, M<T>::abs(p.z) - (T)0.5 * b.z};
  P3<T> qp = {M<T>::fmax(q.x, (T)0), M<T>::fmax(q.y, (T)0), M<T>::fmax(q.z, (T)0)};
  return length(qp) + M<T>::fmin(M<T>::fmax(q.x, M<T>::fmax(q.y, q.z)), (T)0);
}
template <class T> T prim_cylinder(P3<T> p, T h, T r) {  // :13-17
  P2<T> d = {M<T>::abs(length(P2<T>{p.x, p.z})) - r, M<T>::abs(p.y) - h};
  P2<T> dp = {M<T>::fmax(d.x, (T)0), M<T>::fmax(d.y, (T)0)};
  return M<T>::fmin(M<T>::fmax(d.x, d.y), (T)0) + length(dp);
}
template <class T> T prim_capsule(P3<T> p, P3<T> a, P3<T> b, T r) {  // :19-25
  P3<T> pa = p - a, ba = b - a;
  T h = clamp(dot(pa, ba) / dot(ba, ba), (T)0, (T)1);
  return length(pa - ba * h) - r;
}
template <class T> T prim_sphere(P3<T> p, T s) { return length(p) - s; }  // :27-30
template <class T> T prim_torus(P3<T> p, P2<T> t) {                        // :32-36
  P2<T> q = {length(P2<T>{p.x, p.z}) - t.x, p.y};
  return length(q) - t.y;
}
// ---- src/sdf_op.wgsl
template <class T> T op_union(T d1, T d2, T k) {         // :7-11
  T h = clamp((T)0.5 + (T)0.5 * (d2 - d1) / k, (T)0, (T)1);
  return mix(d2, d1, h) - k * h * ((T)1 - h);
}
template <class T> T op_intersection(T d1, T d2, T k) {  // :13-17
  T h = clamp((T)0.5 - (T)0.5 * (d2 - d1) / k, (T)0, (T)1);
  return mix(d2, d1, h) + k * h * ((T)1 - h);
}
template <class T> T op_subtraction(T d1, T d2, T k) {   // :19-23
  T h = clamp((T)0.5 - (T)0.5 * (d2 + d1) / k, (T)0, (T)1);
  return mix(d2, -d1, h) + k * h * ((T)1 - h);
}

// ---- examples/torus.sdf3d
template <class T> T ex_torus(P3<T> p) { return prim_torus(p, P2<T>{K<T>(0.5), K<T>(0.2)}); }

// ---- the arc both letter files define (martin_cube.sdf3d:10-28 letter_r_arc with C, p_key.sdf3d:9-27
// letter_p_arc with KEY_SIZE); S is that module constant
template <class T> T ex_arc(P3<T> p, T ra, T rb, double S) {
  P3<T> pp = {p.y, p.x, p.z};
  pp.x = M<T>::abs(pp.x - K<T>(S * 0.2));
  pp.y = pp.y + K<T>(S * 0.25);
  const T hy = K<T>(S * 0.15);
  // q = pp - clamp(pp, -h, h) with h = (0, S*0.15, 0)
  P3<T> q = {pp.x - clamp(pp.x, K<T>(-0.0), K<T>(0.0)), pp.y - clamp(pp.y, -hy, hy), pp.z - clamp(pp.z, K<T>(-0.0), K<T>(0.0))};
  const P2<T> sc = {(T)1, (T)0};
  T k;
  if ((T)0 > sc.x * q.y) k = dot(P2<T>{q.x, q.y}, sc);
  else k = length(P2<T>{q.x, q.y});
  return M<T>::sqrt(dot(q, q) + ra * ra - (T)2 * ra * k) - rb;
}

// ---- examples/martin_cube.sdf3d
template <class T> struct MartinCube {
  static constexpr double W = 1.0, TH = 0.12, SM = 0.02, C = W;  // :4-8
  static P3<T> v(double x, double y, double z) { return {K<T>(x), K<T>(y), K<T>(z)}; }
  static T seg(P3<T> p, P3<T> a, P3<T> b) { return prim_capsule(p, a, b, K<T>(TH)); }  // :31-33
  static T m(P3<T> p) {                                                                // :36-49
    P3<T> q = {p.y, p.x, p.z};
    q.y = q.y * (T)-1;
    return M<T>::fmin(M<T>::fmin(seg(q, v(-C * 0.45, -C * 0.5, C), v(-C * 0.45, C * 0.5, C)), seg(q, v(-C * 0.45, -C * 0.5, C), v(0, 0, C))),
                      M<T>::fmin(seg(q, v(0, 0, C), v(C * 0.45, -C * 0.5, C)), seg(q, v(C * 0.45, -C * 0.5, C), v(C * 0.45, C * 0.5, C))));
  }
  static T a(P3<T> p) {  // :52-63
    P3<T> q = {p.z, p.y, p.x};
    q = P3<T>{-q.y, -q.x, q.z};
    return M<T>::fmin(M<T>::fmin(seg(q, v(0, -C * 0.5, -C), v(-C * 0.4, C * 0.5, -C)), seg(q, v(0, -C * 0.5, -C), v(C * 0.4, C * 0.5, -C))),
                      seg(q, v(-C * 0.2, C * 0.1, -C), v(C * 0.2, C * 0.1, -C)));
  }
  static T r(P3<T> p) {  // :66-77
    P3<T> q = p;
    q.y = q.y * (T)-1;
    q.x = q.x - K<T>(C * 0.15);
    return M<T>::fmin(ex_arc(P3<T>{q.x, q.z, q.y} - v(0, 0, C), K<T>(C * 0.3), K<T>(TH), C),
                      M<T>::fmin(seg(q, v(-C * 0.4, C, C * 0.5), v(-C * 0.4, C, -C * 0.5)), seg(q, v(-C * 0.15, C, -C * 0.1), v(C * 0.15, C, -C * 0.5))));
  }
  static T t(P3<T> p) {  // :81-86
    return M<T>::fmin(seg(p, v(C, 0, C * 0.5), v(C, 0, -C * 0.5)), seg(p, v(C, -C * 0.4, C * 0.5), v(C, C * 0.4, C * 0.5)));
  }
  static T i(P3<T> p) {  // :90-94
    P3<T> q = {p.y, p.x, p.z};
    q.x = q.x * (T)-1;
    return seg(q, v(-C, 0, C * 0.5), v(-C, 0, -C * 0.5));
  }
  static T n(P3<T> p) {  // :98-109
    P3<T> q = {p.x, p.z, p.y};
    return M<T>::fmin(seg(q, v(-C * 0.4, -C, C * 0.5), v(-C * 0.4, -C, -C * 0.5)),
                      M<T>::fmin(seg(q, v(C * 0.4, -C, C * 0.5), v(C * 0.4, -C, -C * 0.5)), seg(q, v(-C * 0.4, -C, -C * 0.5), v(C * 0.4, -C, C * 0.5))));
  }
  static T sdf(P3<T> p) {  // :111-122
    T cube = op_intersection(prim_box(p, v(W, W, W)), prim_sphere(p, K<T>(W * 1.40)), K<T>(SM));
    T letter = M<T>::fmin(M<T>::fmin(m(p), a(p)), M<T>::fmin(M<T>::fmin(r(p), t(p)), M<T>::fmin(i(p), n(p))));
    return op_subtraction(letter, cube, K<T>(SM));
  }
};

// ---- examples/p_key.sdf3d
template <class T> struct PKey {
  static constexpr double SZ = 15.0, EL = 2.0, PH = 1.5, TH = 1.0;  // :4-7
  static P3<T> v(double x, double y, double z) { return {K<T>(x), K<T>(y), K<T>(z)}; }
  static T seg(P3<T> p, P3<T> a, P3<T> b) { return prim_capsule(p, a, b, K<T>(TH)); }  // :29-31
  static T letter(P3<T> p) {                                                            // :33-41
    P3<T> q = p;
    q.y = q.y * (T)-1;
    q.x = q.x - K<T>(SZ * 0.15);
    return M<T>::fmin(ex_arc(P3<T>{q.x, q.z, q.y}, K<T>(SZ * 0.3), K<T>(TH), SZ), seg(q, v(-SZ * 0.4, 0, SZ * 0.5), v(-SZ * 0.4, 0, -SZ * 0.5)));
  }
  static T sdf(P3<T> p) {  // :43-52
    P3<T> lp = P3<T>{p.x, p.z, p.y} * (T)2.5 - v(0, EL * 2.5 + TH, 0);
    P3<T> cp = P3<T>{p.y, p.z, p.x} - v(0, (EL + PH) * 0.5, 0);
    return op_subtraction(letter(lp), op_union(prim_box(p, v(SZ, SZ, PH)), prim_cylinder(cp, K<T>(EL - PH), K<T>(SZ * 0.35)), K<T>(SZ * 0.15)), (T)0);
  }
};

// ---- examples/mandelmesh.frag:3-28.  GLSL literals are f32; `continue` keeps r = length(z).
template <class T> T ex_mandelbulb(P3<T> pin) {
  const P3<T> p = {pin.x, pin.z, pin.y};
  P3<T> z = p;
  const T power = (T)8;
  T r = 0, theta = 0, phi = 0, dr = 1;
  for (int it = 0; it < 5; ++it) {
    r = length(z);
    if (r > (T)2) continue;
    theta = M<T>::atan(z.y / z.x);
    phi = M<T>::asin(z.z / r);
    dr = M<T>::pow(r, power - (T)1) * dr * power + (T)1;
    r = M<T>::pow(r, power);
    theta = theta * power;
    phi = phi * power;
    const P3<T> dir = {M<T>::cos(theta) * M<T>::cos(phi), M<T>::sin(theta) * M<T>::cos(phi), M<T>::sin(phi)};
    z = dir * r + p;
  }
  return (T)0.5 * M<T>::log(r) * r / dr - K<T>(0.003);
}

// ---- src/shadertoy.rs:411-442 (the GLSL of test_naga; iTime = 0)
template <class T> T ex_naga_sphere(P3<T> p) {
  const T sphere = length(p - P3<T>{0, 0, 0}) - (T)1;
  const T t = 0;
  const T disp = M<T>::sin((T)5 * p.x) * M<T>::sin((T)5 * p.y) * M<T>::sin((T)5 * p.z) * (T)0.25 * M<T>::sin((T)2 * t);
  return sphere + disp;
}

template <class T> T eval_sdf(int id, P3<T> p) {
  switch (id) {
    case 0: return ex_torus(p);
    case 1: return MartinCube<T>::sdf(p);
    case 2: return PKey<T>::sdf(p);
    case 3: return ex_mandelbulb(p);
    case 4: return ex_naga_sphere(p);
  }
  return (T)0;
}

struct Rec {
  uint64_t key;      // x | y<<16 | label<<32 (mesh.rs:224-226)
  double pos[3];     // mean-of-crossings position (valid if active)
  double min_abs;    // min |corner value| over the 8 corners, in SDF units
  double corner[8];  // the corner values, reference order 000,100,010,110,001,101,011,111
  uint8_t active;    // the reference's rule gives this cell a vertex (count >= 2)
  uint8_t nibble;    // cell_sign_changes (dualcontour.wgsl:57-69)
  uint8_t pad[6];
};

struct Run {
  std::vector<Rec> recs;
  uint64_t n_cells = 0;
};

template <class T>
void run_rows(int sdf, const uint32_t res[3], const float bmin[3], const float bmax[3], uint32_t z, uint32_t label, double list_below,
              int y0, int y1, std::vector<Rec>& out) {
  float size[3];
  for (int a = 0; a < 3; ++a) size[a] = (bmax[a] - bmin[a]) / (float)(res[a] - 1u);  // dualcontour.wgsl:23-24, f32 as in the shader
  const float zmin = bmin[2] + size[2] * (float)z, zmax = zmin + size[2];
  static const int ea[12] = {0, 2, 1, 3, 0, 4, 1, 5, 0, 4, 2, 6};  // :99-114, Z edges, Y edges, X edges
  static const int eb[12] = {4, 6, 5, 7, 2, 6, 3, 7, 1, 5, 3, 7};
  static const int ax[12] = {2, 2, 2, 2, 1, 1, 1, 1, 0, 0, 0, 0};
  for (int y = y0; y < y1; ++y) {
    const float ymin = bmin[1] + size[1] * (float)y, ymax = ymin + size[1];
    for (uint32_t x = 0; x < res[0]; ++x) {
      const float xmin = bmin[0] + size[0] * (float)x, xmax = xmin + size[0];
      T d[8];
      double mabs = INFINITY;
      for (int c = 0; c < 8; ++c) {
        d[c] = eval_sdf<T>(sdf, P3<T>{(T)((c & 1) ? xmax : xmin), (T)((c & 2) ? ymax : ymin), (T)((c & 4) ? zmax : zmin)});
        const double a = std::fabs((double)d[c]);
        if (!(a >= mabs)) mabs = a;  // NaN counts as 0 distance: always listed
        if (a != a) mabs = 0.0;
      }
      T avg[3] = {0, 0, 0};
      T count = 0;
      for (int e = 0; e < 12; ++e) {
        const T v0 = d[ea[e]], v1 = d[eb[e]];
        if ((v0 > (T)0) != (v1 > (T)0)) {
          T ch[3] = {(T)((ea[e] & 1) ? 1 : 0), (T)((ea[e] & 2) ? 1 : 0), (T)((ea[e] & 4) ? 1 : 0)};
          ch[ax[e]] = ((T)0 - v0) / (v1 - v0);  // _cell_adapt :72-74
          if (ch[0] > (T)0 || ch[1] > (T)0 || ch[2] > (T)0) { avg[0] += ch[0]; avg[1] += ch[1]; avg[2] += ch[2]; count += (T)1; }
        }
      }
      const bool active = !(count <= (T)1);
      if (!active && !(mabs < list_below)) continue;
      Rec r;
      memset(&r, 0, sizeof r);
      r.key = (uint64_t)x | ((uint64_t)y << 16) | ((uint64_t)label << 32);
      r.active = active ? 1 : 0;
      r.min_abs = mabs;
      for (int c = 0; c < 8; ++c) r.corner[c] = (double)d[c];
      if (active) {
        const T cmin[3] = {(T)xmin, (T)ymin, (T)zmin}, cmax[3] = {(T)xmax, (T)ymax, (T)zmax};
        for (int a = 0; a < 3; ++a) r.pos[a] = (double)(cmin[a] + (cmax[a] - cmin[a]) * avg[a] / count);  // :130
      }
      r.nibble = (uint8_t)((d[1] > (T)0 ? 1 : 0) | (d[2] > (T)0 ? 2 : 0) | (d[4] > (T)0 ? 4 : 0) | (d[0] > (T)0 ? 8 : 0));
      out.push_back(r);
    }
  }
}

}  // namespace

extern "C" {

// Cells of true slices [z_begin, z_end), label = z + label_add.  Returned: every cell that is active
// under the reference's rule evaluated in T, plus every cell with a corner of |value| < list_below
// (SDF units).  precision: 64 = double + libm f64, 32 = float + libm f32.
void* indep_run(int sdf, const uint32_t res[3], const float bmin[3], const float bmax[3], uint32_t z_begin, uint32_t z_end,
                uint32_t label_add, double list_below, int precision, int threads) {
  Run* run = new Run();
  if (threads <= 0) { threads = (int)std::thread::hardware_concurrency(); if (threads <= 0) threads = 1; }
  const int ny = (int)res[1];
  for (uint32_t z = z_begin; z < z_end; ++z) {
    std::vector<std::vector<Rec>> rows((size_t)ny);
    std::atomic<int> next{0};
    auto work = [&]() {
      for (;;) {
        const int y0 = next.fetch_add(2);
        if (y0 >= ny) break;
        for (int y = y0; y < std::min(ny, y0 + 2); ++y) {
          if (precision == 32) run_rows<float>(sdf, res, bmin, bmax, z, z + label_add, list_below, y, y + 1, rows[(size_t)y]);
          else run_rows<double>(sdf, res, bmin, bmax, z, z + label_add, list_below, y, y + 1, rows[(size_t)y]);
        }
      }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(work);
    work();
    for (auto& th : pool) th.join();
    for (int y = 0; y < ny; ++y) run->recs.insert(run->recs.end(), rows[(size_t)y].begin(), rows[(size_t)y].end());
    run->n_cells += (uint64_t)res[0] * res[1];
  }
  return run;
}
uint64_t indep_count(void* h) { return ((Run*)h)->recs.size(); }
// keys u64[n], active u8[n], nibble u8[n], pos f64[n][3], min_abs f64[n], corners f64[n][8]
void indep_copy(void* h, uint64_t* keys, uint8_t* active, uint8_t* nibble, double* pos, double* min_abs, double* corners) {
  const Run* run = (Run*)h;
  for (size_t i = 0; i < run->recs.size(); ++i) {
    const Rec& r = run->recs[i];
    keys[i] = r.key; active[i] = r.active; nibble[i] = r.nibble; min_abs[i] = r.min_abs;
    for (int a = 0; a < 3; ++a) pos[3 * i + a] = r.pos[a];
    if (corners) for (int c = 0; c < 8; ++c) corners[8 * i + c] = r.corner[c];
  }
}
void indep_free(void* h) { delete (Run*)h; }
// one SDF value (diagnostics): precision as above
double indep_eval(int sdf, double x, double y, double z, int precision) {
  if (precision == 32) return (double)eval_sdf<float>(sdf, P3<float>{(float)x, (float)y, (float)z});
  return eval_sdf<double>(sdf, P3<double>{x, y, z});
}

}  // extern "C"
