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
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "common.h"

using s2m_internal::fail;

namespace {

// appends Rust `{}` formatting of f to out
void rust_f32(float f, std::string& out) {
  if (f != f) { out += "NaN"; return; }
  if (std::isinf(f)) { out += f > 0 ? "inf" : "-inf"; return; }
  if (f == 0.0f) { out += std::signbit(f) ? "-0" : "0"; return; }
  char buf[48];
  auto r = std::to_chars(buf, buf + sizeof buf, f, std::chars_format::scientific);  // shortest round-trip
  // buf = [-]d[.ddd]e[+-]XX
  const char* p = buf;
  if (*p == '-') { out += '-'; ++p; }
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
    out += "0.";
    out.append((size_t)(-ex - 1), '0');
    out.append(digits, (size_t)nd);
  } else if (ex + 1 >= nd) {
    out.append(digits, (size_t)nd);
    out.append((size_t)(ex + 1 - nd), '0');
  } else {
    out.append(digits, (size_t)(ex + 1));
    out += '.';
    out.append(digits + ex + 1, (size_t)(nd - ex - 1));
  }
}

// One or more z-slab results, in z order, addressed by global vertex index.
struct Part {
  s2m_result_info i;
  int64_t base;   // global index of this part's first own vertex
};
inline uint64_t quad_index(const s2m_result_info& m, uint64_t i) { return m.quads ? m.quads[i] : (uint64_t)m.quads32[i]; }

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

// run fn(begin, end, out) over [0, n) in ordered chunks, formatting chunks on several threads and
// writing them in order
template <class Fn>
int parallel_write(FILE* f, uint64_t n, uint64_t chunk, Fn fn) {
  const unsigned hw = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
  const uint64_t n_chunks = (n + chunk - 1) / chunk;
  for (uint64_t base = 0; base < n_chunks; base += hw) {
    const unsigned cnt = (unsigned)std::min<uint64_t>(hw, n_chunks - base);
    std::vector<std::string> bufs(cnt);
    std::vector<std::thread> th;
    for (unsigned t = 0; t < cnt; ++t)
      th.emplace_back([&, t] {
        const uint64_t b = (base + t) * chunk, e = std::min(n, b + chunk);
        bufs[t].reserve((size_t)(e - b) * 96);
        fn(b, e, bufs[t]);
      });
    for (auto& x : th) x.join();
    for (unsigned t = 0; t < cnt; ++t)
      if (fwrite(bufs[t].data(), 1, bufs[t].size(), f) != bufs[t].size()) return fail(S2M_ERR_IO, "short write");
  }
  return S2M_OK;
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
    st = parallel_write(f, m.n_quads, 1u << 15, [&](uint64_t b, uint64_t e, std::string& out) {
      for (uint64_t q = b; q < e; ++q)
        for (int t = 0; t < 2; ++t) {
          uint64_t tri[3];
          tri_of_quad(m, q, t, tri);
          const float* p0 = v.pos(tri[0], k);
          const float* p1 = v.pos(tri[1], k);
          const float* p2 = v.pos(tri[2], k);
          float n[3];
          tri_normal(p0, p1, p2, n);
          out += "facet normal ";
          rust_f32(n[0], out); out += ' '; rust_f32(n[1], out); out += ' '; rust_f32(n[2], out);
          out += "\n\touter loop\n";
          for (const float* p : {p0, p1, p2}) {
            out += "\t\tvertex ";
            rust_f32(p[0], out); out += ' '; rust_f32(p[1], out); out += ' '; rust_f32(p[2], out);
            out += '\n';
          }
          out += "\tendloop\nendfacet\n";
        }
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
    st = parallel_write(f, m.n_vertices, 1u << 16, [&](uint64_t b, uint64_t e, std::string& out) {
      for (uint64_t i = b; i < e; ++i) {
        const float* p = m.positions + 3 * i;
        const float* n = m.normals + 3 * i;
        rust_f32(p[0], out); out += ' '; rust_f32(p[1], out); out += ' '; rust_f32(p[2], out); out += ' ';
        rust_f32(n[0], out); out += ' '; rust_f32(n[1], out); out += ' '; rust_f32(n[2], out); out += '\n';
      }
    });
  }
  for (size_t k = 0; k < v.parts.size() && st == S2M_OK; ++k) {
    const s2m_result_info& m = v.parts[k].i;
    st = parallel_write(f, m.n_quads, 1u << 16, [&](uint64_t b, uint64_t e, std::string& out) {
      char buf[128];
      for (uint64_t q = b; q < e; ++q)
        for (int t = 0; t < 2; ++t) {
          uint64_t tri[3];
          tri_of_quad(m, q, t, tri);
          int n = snprintf(buf, sizeof buf, "3 %u %u %u\n", (unsigned)tri[0], (unsigned)tri[1], (unsigned)tri[2]);  // Triangle<u32>
          out.append(buf, (size_t)n);
        }
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
    st = parallel_write(f, m.n_quads, 1u << 16, [&](uint64_t b, uint64_t e, std::string& out) {
      out.resize((size_t)(e - b) * 100);
      char* o = &out[0];
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
