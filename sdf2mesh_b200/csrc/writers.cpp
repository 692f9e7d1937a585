// writers.cpp -- mesh file writers, byte-compatible with the reference's text formats:
//   ASCII STL  /root/reference/src/mesh.rs:8-48, :155-180   (+ Triangle::normal, /root/reference/src/lib.rs:181-185)
//   ASCII PLY  /root/reference/src/mesh.rs:50-141, :198-210
//   dispatch   /root/reference/src/mesh.rs:182-196 (extension, case-insensitive; unknown -> logged, Ok)
// Numbers are printed like Rust's `{}` for f32: the shortest decimal that round-trips, never in
// exponent form, integers without ".0", "NaN", "inf", "-inf", "-0".
// Binary STL (50 bytes / triangle) is the additional fast format north_star names.
// Formatting is spread over host threads (a 2048^3 mandelbulb is ~17 M triangles = ~4 GB of text).
#include <algorithm>
#include <atomic>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <system_error>
#include <thread>
#include <vector>

#include "common.h"

using s2m_internal::fail;

namespace {

// Rust `{}` formatting of f at o (at most kMaxF32Chars bytes); returns the end
constexpr size_t kMaxF32Chars = 64;  // "-0." + 44 zeros + 9 digits
inline char* put(char* o, const char* text, size_t n) { memcpy(o, text, n); return o + n; }
template <size_t N>
inline char* put(char* o, const char (&text)[N]) { return put(o, text, N - 1); }
char* rust_f32(float f, char* o) {
  if (f != f) return put(o, "NaN");
  if (std::isinf(f)) return f > 0 ? put(o, "inf") : put(o, "-inf");
  if (f == 0.0f) return std::signbit(f) ? put(o, "-0") : put(o, "0");
  char buf[48];
  auto r = std::to_chars(buf, buf + sizeof buf, f, std::chars_format::scientific);  // shortest round-trip
  // buf = [-]d[.ddd]e[+-]XX
  const char* p = buf;
  if (*p == '-') { *o++ = '-'; ++p; }
  char digits[16];
  int nd = 0;
  while (p < r.ptr && *p != 'e') { if (*p != '.') digits[nd++] = *p; ++p; }
  int ex = 0;
  if (p < r.ptr) {
    ++p;
    bool neg = false;
    if (*p == '-') { neg = true; ++p; } else if (*p == '+') ++p;
    while (p < r.ptr) ex = ex * 10 + (*p++ - '0');
    if (neg) ex = -ex;
  }
  while (nd > 1 && digits[nd - 1] == '0') --nd;
  if (ex < 0) {
    *o++ = '0'; *o++ = '.';
    memset(o, '0', (size_t)(-ex - 1)); o += -ex - 1;
    return put(o, digits, (size_t)nd);
  }
  if (ex + 1 >= nd) {
    o = put(o, digits, (size_t)nd);
    memset(o, '0', (size_t)(ex + 1 - nd));
    return o + (ex + 1 - nd);
  }
  o = put(o, digits, (size_t)(ex + 1));
  *o++ = '.';
  return put(o, digits + ex + 1, (size_t)(nd - ex - 1));
}
inline char* rust_f32x3(const float* v, char* o, char last) {
  o = rust_f32(v[0], o); *o++ = ' ';
  o = rust_f32(v[1], o); *o++ = ' ';
  o = rust_f32(v[2], o); *o++ = last;
  return o;
}
inline char* put_u32(char* o, uint32_t v) { return std::to_chars(o, o + 10, v).ptr; }

// One or more z-slab results, in z order, addressed by global vertex index.
struct Part {
  s2m_result_info i;
  int64_t base;   // global index of this part's first own vertex
};
// global vertex index of quad entry i: value + quad_index_add, wrapping in the index width (S2M_MESH_RELATIVE_QUADS)
inline uint64_t quad_index(const s2m_result_info& m, uint64_t i) {
  return m.quads ? m.quads[i] + (uint64_t)m.quad_index_add : (uint64_t)(uint32_t)(m.quads32[i] + (uint32_t)m.quad_index_add);
}

struct Parts {
  std::vector<Part> parts;
  uint64_t n_vertices = 0, n_quads = 0;
  // position of global vertex g as seen from part `hint` (its own vertices, its halo, or another part)
  const float* pos(uint64_t g, size_t hint) const {
    const Part& h = parts[hint];
    const int64_t gi = (int64_t)g;
    if (gi >= h.base && gi < h.base + (int64_t)h.i.n_vertices) return h.i.positions + 3 * (gi - h.base);
    const int64_t hb = h.base - (int64_t)h.i.n_halo_vertices;
    if (gi >= hb && gi < h.base && h.i.halo_positions) return h.i.halo_positions + 3 * (gi - hb);
    for (const Part& p : parts)
      if (gi >= p.base && gi < p.base + (int64_t)p.i.n_vertices) return p.i.positions + 3 * (gi - p.base);
    return nullptr;
  }
};

int get_parts(const s2m_result_info* infos, int n, Parts* out, bool whole_mesh) {
  if (!infos || n <= 0) return fail(S2M_ERR_INVALID_ARG, "no results to write");
  for (int k = 0; k < n; ++k) {
    Part p;
    p.i = infos[k];
    if (p.i.n_quads && !p.i.quads && !p.i.quads32) return fail(S2M_ERR_STATE, "s2m_mesh_finish has not been called");
    if (p.i.n_vertices && !p.i.positions) return fail(S2M_ERR_INVALID_ARG, "positions is NULL");
    if (whole_mesh && p.i.n_vertices && !p.i.normals) return fail(S2M_ERR_INVALID_ARG, "normals is NULL (PLY writes them)");
    p.base = p.i.global_vertex_base;
    if (k > 0 && p.base != out->parts.back().base + (int64_t)out->parts.back().i.n_vertices)
      return fail(S2M_ERR_STATE, "parts must be consecutive z-slabs with consecutive global vertex bases");
    out->parts.push_back(p);
    out->n_vertices += p.i.n_vertices;
    out->n_quads += p.i.n_quads;
  }
  if (whole_mesh && out->parts.front().base != 0) return fail(S2M_ERR_STATE, "this format needs the whole mesh (first part must start at vertex 0)");
  // every quad index must resolve
  for (size_t k = 0; k < out->parts.size(); ++k) {
    const Part& p = out->parts[k];
    for (uint64_t q = 0; q < 4 * p.i.n_quads; ++q)
      if (!out->pos(quad_index(p.i, q), k)) return fail(S2M_ERR_STATE, "a quad refers to a vertex outside the given parts");
  }
  return S2M_OK;
}
int get_parts(const s2m_result* const* rs, int n, Parts* out, bool whole_mesh) {
  if (!rs || n <= 0) return fail(S2M_ERR_INVALID_ARG, "no results to write");
  std::vector<s2m_result_info> infos((size_t)n);
  for (int k = 0; k < n; ++k) {
    if (!rs[k]) return fail(S2M_ERR_INVALID_ARG, "result is NULL");
    int st = s2m_result_get(rs[k], &infos[(size_t)k]);
    if (st) return st;
  }
  return get_parts(infos.data(), n, out, whole_mesh);
}

// Runs `char* fn(begin, end, char* out)` over [0, n) in chunks of `chunk` items (at most `max_item_bytes`
// of output each) and writes the chunks to f in order.  Worker threads take chunk numbers from a
// counter and format into a ring of reusable buffers; the calling thread writes buffer after buffer as
// they become ready, so formatting and fwrite overlap and no memory is touched for the first time after
// the ring has been filled once.
template <class Fn>
int parallel_write(FILE* f, uint64_t n, uint64_t chunk, size_t max_item_bytes, Fn fn) {
  if (n == 0) return S2M_OK;
  if (const char* e = getenv("S2M_WRITER_CHUNK")) chunk = (uint64_t)std::max(1, atoi(e));  // tests: many chunks from a small mesh
  const uint64_t n_chunks = (n + chunk - 1) / chunk;
  unsigned hw = std::max(1u, std::min(64u, std::thread::hardware_concurrency()));
  if (const char* e = getenv("S2M_WRITER_THREADS")) hw = (unsigned)std::max(0, std::min(256, atoi(e)));
  const unsigned workers = (unsigned)std::min<uint64_t>(hw, n_chunks);
  auto serial_from = [&](uint64_t first) {  // no worker threads (S2M_WRITER_THREADS=0, or the process may not create any): format and write in turn
    std::unique_ptr<char[]> buf(new char[(size_t)chunk * max_item_bytes]);
    for (uint64_t c = first; c < n_chunks; ++c) {
      const uint64_t b = c * chunk, e = std::min(n, b + chunk);
      const size_t len = (size_t)(fn(b, e, buf.get()) - buf.get());
      if (fwrite(buf.get(), 1, len, f) != len) return fail(S2M_ERR_IO, std::string("short write: ") + strerror(errno));
    }
    return (int)S2M_OK;
  };
  if (workers == 0) return serial_from(0);
  const unsigned n_slots = 2 * workers;
  struct Slot {
    std::unique_ptr<char[]> buf;
    size_t len = 0;
    uint64_t holds = ~0ull;   // chunk number whose text is in buf
  };
  std::vector<Slot> slots(n_slots);
  for (Slot& sl : slots) sl.buf.reset(new char[(size_t)chunk * max_item_bytes]);  // untouched pages cost nothing
  std::mutex mu;
  std::condition_variable cv_ready, cv_free;
  uint64_t next_chunk = 0, written = 0;   // guarded by mu; chunk c may use its slot once c < written + n_slots
  bool stop = false;
  auto work = [&] {
    for (;;) {
      uint64_t c;
      {
        std::unique_lock<std::mutex> lk(mu);
        if (stop || next_chunk >= n_chunks) return;
        c = next_chunk++;
        cv_free.wait(lk, [&] { return stop || c < written + n_slots; });
        if (stop) return;
      }
      Slot& sl = slots[c % n_slots];
      const uint64_t b = c * chunk, e = std::min(n, b + chunk);
      sl.len = (size_t)(fn(b, e, sl.buf.get()) - sl.buf.get());
      {
        std::lock_guard<std::mutex> lk(mu);
        sl.holds = c;
      }
      cv_ready.notify_one();
    }
  };
  std::vector<std::thread> th;
  try {
    for (unsigned t = 0; t < workers; ++t) th.emplace_back(work);
  } catch (const std::system_error&) {   // thread limit of the process: carry on with the threads there are
    if (th.empty()) return serial_from(0);
  }
  int st = S2M_OK;
  for (uint64_t c = 0; c < n_chunks; ++c) {
    Slot& sl = slots[c % n_slots];
    {
      std::unique_lock<std::mutex> lk(mu);
      cv_ready.wait(lk, [&] { return sl.holds == c; });
    }
    const bool ok = fwrite(sl.buf.get(), 1, sl.len, f) == sl.len;
    {
      std::lock_guard<std::mutex> lk(mu);
      written = c + 1;
      if (!ok) stop = true;
    }
    cv_free.notify_all();
    if (!ok) { st = fail(S2M_ERR_IO, std::string("short write: ") + strerror(errno)); break; }
  }
  for (auto& x : th) x.join();
  return st;
}

inline void tri_of_quad(const s2m_result_info& m, uint64_t quad, int t, uint64_t tri[3]) {  // lib.rs:199-204
  const uint64_t q[4] = {quad_index(m, 4 * quad), quad_index(m, 4 * quad + 1), quad_index(m, 4 * quad + 2), quad_index(m, 4 * quad + 3)};
  if (t == 0) { tri[0] = q[2]; tri[1] = q[1]; tri[2] = q[0]; }
  else { tri[0] = q[0]; tri[1] = q[3]; tri[2] = q[2]; }
}
inline void tri_normal(const float* p0, const float* p1, const float* p2, float n[3]) {  // lib.rs:181-185
  const float a[3] = {p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2]};
  const float b[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]};
  n[0] = a[1] * b[2] - a[2] * b[1];
  n[1] = a[2] * b[0] - a[0] * b[2];
  n[2] = a[0] * b[1] - a[1] * b[0];
}

int write_stl_ascii(const Parts& v, const char* path) {
  FILE* f = fopen(path, "wb");
  if (!f) return fail(S2M_ERR_IO, std::string("cannot create ") + path + ": " + strerror(errno));
  fputs("solid\n", f);
  int st = S2M_OK;
  for (size_t k = 0; k < v.parts.size() && st == S2M_OK; ++k) {
    const s2m_result_info& m = v.parts[k].i;
    // per quad: 2 facets of 4 float triples and 70 bytes of keywords
    st = parallel_write(f, m.n_quads, 1u << 13, 2 * (12 * (kMaxF32Chars + 1) + 70), [&](uint64_t b, uint64_t e, char* o) {
      for (uint64_t q = b; q < e; ++q)
        for (int t = 0; t < 2; ++t) {
          uint64_t tri[3];
          tri_of_quad(m, q, t, tri);
          const float* p0 = v.pos(tri[0], k);
          const float* p1 = v.pos(tri[1], k);
          const float* p2 = v.pos(tri[2], k);
          float n[3];
          tri_normal(p0, p1, p2, n);
          o = put(o, "facet normal ");
          o = rust_f32x3(n, o, '\n');
          o = put(o, "\touter loop\n");
          for (const float* p : {p0, p1, p2}) {
            o = put(o, "\t\tvertex ");
            o = rust_f32x3(p, o, '\n');
          }
          o = put(o, "\tendloop\nendfacet\n");
        }
      return o;
    });
  }
  fputs("endsolid\n", f);
  if (fclose(f) != 0 && st == S2M_OK) st = fail(S2M_ERR_IO, std::string("close failed for ") + path);
  return st;
}

int write_ply_ascii(const Parts& v, const char* path) {
  FILE* f = fopen(path, "wb");
  if (!f) return fail(S2M_ERR_IO, std::string("cannot create ") + path + ": " + strerror(errno));
  fprintf(f, "ply\nformat ascii 1.0\ncomment written by rust-sdf\n");
  fprintf(f, "element vertex %llu\nproperty float x\nproperty float y\nproperty float z\nproperty float nx\nproperty float ny\nproperty float nz\n",
          (unsigned long long)v.n_vertices);
  fprintf(f, "element face %llu\nproperty list uchar int vertex_index\nend_header\n", (unsigned long long)(2 * v.n_quads));
  int st = S2M_OK;
  for (size_t k = 0; k < v.parts.size() && st == S2M_OK; ++k) {
    const s2m_result_info& m = v.parts[k].i;
    st = parallel_write(f, m.n_vertices, 1u << 14, 6 * (kMaxF32Chars + 1), [&](uint64_t b, uint64_t e, char* o) {
      for (uint64_t i = b; i < e; ++i) {
        o = rust_f32x3(m.positions + 3 * i, o, ' ');
        o = rust_f32x3(m.normals + 3 * i, o, '\n');
      }
      return o;
    });
  }
  for (size_t k = 0; k < v.parts.size() && st == S2M_OK; ++k) {
    const s2m_result_info& m = v.parts[k].i;
    st = parallel_write(f, m.n_quads, 1u << 15, 2 * (3 * 11 + 2), [&](uint64_t b, uint64_t e, char* o) {
      for (uint64_t q = b; q < e; ++q)
        for (int t = 0; t < 2; ++t) {
          uint64_t tri[3];
          tri_of_quad(m, q, t, tri);
          *o++ = '3';
          for (int c = 0; c < 3; ++c) { *o++ = ' '; o = put_u32(o, (uint32_t)tri[c]); }  // Triangle<u32>
          *o++ = '\n';
        }
      return o;
    });
  }
  if (fclose(f) != 0 && st == S2M_OK) st = fail(S2M_ERR_IO, std::string("close failed for ") + path);
  return st;
}

int write_stl_binary(const Parts& v, const char* path) {
  if (2 * v.n_quads > 0xffffffffull) return fail(S2M_ERR_UNSUPPORTED, "binary STL holds at most 2^32-1 triangles");
  FILE* f = fopen(path, "wb");
  if (!f) return fail(S2M_ERR_IO, std::string("cannot create ") + path + ": " + strerror(errno));
  char header[80];
  memset(header, 0, sizeof header);
  snprintf(header, sizeof header, "sdf2mesh_b200 binary STL");
  const uint32_t ntri = (uint32_t)(2 * v.n_quads);
  fwrite(header, 1, 80, f);
  fwrite(&ntri, 4, 1, f);
  int st = S2M_OK;
  for (size_t k = 0; k < v.parts.size() && st == S2M_OK; ++k) {
    const s2m_result_info& m = v.parts[k].i;
    st = parallel_write(f, m.n_quads, 1u << 16, 100, [&](uint64_t b, uint64_t e, char* o) {
      for (uint64_t q = b; q < e; ++q)
        for (int t = 0; t < 2; ++t) {
          uint64_t tri[3];
          tri_of_quad(m, q, t, tri);
          const float* p[3] = {v.pos(tri[0], k), v.pos(tri[1], k), v.pos(tri[2], k)};
          float n[3];
          tri_normal(p[0], p[1], p[2], n);
          const float len = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
          if (len > 0.0f) { n[0] /= len; n[1] /= len; n[2] /= len; } else { n[0] = n[1] = n[2] = 0.0f; }
          memcpy(o, n, 12); o += 12;
          for (int c = 0; c < 3; ++c) { memcpy(o, p[c], 12); o += 12; }
          o[0] = o[1] = 0; o += 2;
        }
      return o;
    });
  }
  if (fclose(f) != 0 && st == S2M_OK) st = fail(S2M_ERR_IO, std::string("close failed for ") + path);
  return st;
}

std::string upper_extension(const char* path) {  // Path::extension().to_ascii_uppercase() (mesh.rs:183-188)
  std::string p(path), ext;
  size_t slash = p.find_last_of('/');
  size_t dot = p.find_last_of('.');
  if (dot != std::string::npos && (slash == std::string::npos || dot > slash) && dot != (slash == std::string::npos ? 0 : slash + 1)) ext = p.substr(dot + 1);
  for (char& c : ext) c = (char)toupper((unsigned char)c);
  return ext;
}

template <class Src>
int write_parts(const Src* rs, int n, const char* path, int binary_stl) {
  if (!path) return fail(S2M_ERR_INVALID_ARG, "path is NULL");
  const std::string ext = upper_extension(path);
  if (ext != "STL" && ext != "PLY") {
    fprintf(stderr, "ERROR Unknown file extension: %s\n", ext.c_str());  // mesh.rs:193; the reference still returns Ok
    return S2M_OK;
  }
  Parts v;
  int st = get_parts(rs, n, &v, ext == "PLY");
  if (st) return st;
  if (ext == "PLY") return write_ply_ascii(v, path);
  return binary_stl ? write_stl_binary(v, path) : write_stl_ascii(v, path);
}

}  // namespace

extern "C" int s2m_result_write_mesh(const s2m_result* r, const char* path) { return write_parts(&r, 1, path, 0); }

extern "C" int s2m_result_write_stl_binary(const s2m_result* r, const char* path) {
  if (!path) return fail(S2M_ERR_INVALID_ARG, "path is NULL");
  Parts v;
  int st = get_parts(&r, 1, &v, false);
  if (st) return st;
  return write_stl_binary(v, path);
}

extern "C" int s2m_write_mesh_parts(const s2m_result* const* parts, int n_parts, const char* path, int binary_stl) {
  return write_parts(parts, n_parts, path, binary_stl);
}

extern "C" int s2m_write_mesh_arrays(const s2m_result_info* parts, int n_parts, const char* path, int binary_stl) {
  return write_parts(parts, n_parts, path, binary_stl);
}
