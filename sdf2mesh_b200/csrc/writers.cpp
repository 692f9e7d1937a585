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

struct MeshView {
  s2m_result_info i;
  bool ok = false;
};

int get_view(const s2m_result* r, MeshView* v) {
  if (!r) return fail(S2M_ERR_INVALID_ARG, "result is NULL");
  int st = s2m_result_get(r, &v->i);
  if (st) return st;
  if (v->i.n_halo_vertices != 0) return fail(S2M_ERR_STATE, "mesh writers need a single-slab result (this one is a z-slab with a halo)");
  if (v->i.n_quads && !v->i.quads) return fail(S2M_ERR_STATE, "s2m_mesh_finish has not been called");
  return S2M_OK;
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

inline void tri_of_quad(const uint64_t* q, int t, uint64_t tri[3]) {  // lib.rs:199-204
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

int write_stl_ascii(const MeshView& v, const char* path) {
  FILE* f = fopen(path, "wb");
  if (!f) return fail(S2M_ERR_IO, std::string("cannot create ") + path + ": " + strerror(errno));
  fputs("solid\n", f);
  const s2m_result_info& m = v.i;
  int st = parallel_write(f, m.n_quads, 1u << 15, [&](uint64_t b, uint64_t e, std::string& out) {
    for (uint64_t q = b; q < e; ++q)
      for (int t = 0; t < 2; ++t) {
        uint64_t tri[3];
        tri_of_quad(m.quads + 4 * q, t, tri);
        const float* p0 = m.positions + 3 * tri[0];
        const float* p1 = m.positions + 3 * tri[1];
        const float* p2 = m.positions + 3 * tri[2];
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
  fputs("endsolid\n", f);
  if (fclose(f) != 0 && st == S2M_OK) st = fail(S2M_ERR_IO, std::string("close failed for ") + path);
  return st;
}

int write_ply_ascii(const MeshView& v, const char* path) {
  FILE* f = fopen(path, "wb");
  if (!f) return fail(S2M_ERR_IO, std::string("cannot create ") + path + ": " + strerror(errno));
  const s2m_result_info& m = v.i;
  fprintf(f, "ply\nformat ascii 1.0\ncomment written by rust-sdf\n");
  fprintf(f, "element vertex %llu\nproperty float x\nproperty float y\nproperty float z\nproperty float nx\nproperty float ny\nproperty float nz\n",
          (unsigned long long)m.n_vertices);
  fprintf(f, "element face %llu\nproperty list uchar int vertex_index\nend_header\n", (unsigned long long)(2 * m.n_quads));
  int st = parallel_write(f, m.n_vertices, 1u << 16, [&](uint64_t b, uint64_t e, std::string& out) {
    for (uint64_t i = b; i < e; ++i) {
      const float* p = m.positions + 3 * i;
      const float* n = m.normals + 3 * i;
      rust_f32(p[0], out); out += ' '; rust_f32(p[1], out); out += ' '; rust_f32(p[2], out); out += ' ';
      rust_f32(n[0], out); out += ' '; rust_f32(n[1], out); out += ' '; rust_f32(n[2], out); out += '\n';
    }
  });
  if (st == S2M_OK)
    st = parallel_write(f, m.n_quads, 1u << 16, [&](uint64_t b, uint64_t e, std::string& out) {
      char buf[128];
      for (uint64_t q = b; q < e; ++q)
        for (int t = 0; t < 2; ++t) {
          uint64_t tri[3];
          tri_of_quad(m.quads + 4 * q, t, tri);
          int n = snprintf(buf, sizeof buf, "3 %u %u %u\n", (unsigned)tri[0], (unsigned)tri[1], (unsigned)tri[2]);  // Triangle<u32>
          out.append(buf, (size_t)n);
        }
    });
  if (fclose(f) != 0 && st == S2M_OK) st = fail(S2M_ERR_IO, std::string("close failed for ") + path);
  return st;
}

}  // namespace

extern "C" int s2m_result_write_mesh(const s2m_result* r, const char* path) {
  if (!path) return fail(S2M_ERR_INVALID_ARG, "path is NULL");
  MeshView v;
  int st = get_view(r, &v);
  if (st) return st;
  for (uint64_t i = 0; i < 4 * v.i.n_quads; ++i)
    if (v.i.quads[i] >= v.i.n_vertices) return fail(S2M_ERR_STATE, "quad index outside this result (a non-zero global vertex base was used)");
  std::string p(path), ext;
  size_t slash = p.find_last_of('/');
  size_t dot = p.find_last_of('.');
  if (dot != std::string::npos && (slash == std::string::npos || dot > slash) && dot != (slash == std::string::npos ? 0 : slash + 1)) ext = p.substr(dot + 1);
  for (char& c : ext) c = (char)toupper((unsigned char)c);
  if (ext == "STL") return write_stl_ascii(v, path);
  if (ext == "PLY") return write_ply_ascii(v, path);
  fprintf(stderr, "ERROR Unknown file extension: %s\n", ext.c_str());  // mesh.rs:193; the reference still returns Ok
  return S2M_OK;
}

extern "C" int s2m_result_write_stl_binary(const s2m_result* r, const char* path) {
  if (!path) return fail(S2M_ERR_INVALID_ARG, "path is NULL");
  MeshView v;
  int st = get_view(r, &v);
  if (st) return st;
  const s2m_result_info& m = v.i;
  if (2 * m.n_quads > 0xffffffffull) return fail(S2M_ERR_UNSUPPORTED, "binary STL holds at most 2^32-1 triangles");
  FILE* f = fopen(path, "wb");
  if (!f) return fail(S2M_ERR_IO, std::string("cannot create ") + path + ": " + strerror(errno));
  char header[80];
  memset(header, 0, sizeof header);
  snprintf(header, sizeof header, "sdf2mesh_b200 binary STL");
  const uint32_t ntri = (uint32_t)(2 * m.n_quads);
  fwrite(header, 1, 80, f);
  fwrite(&ntri, 4, 1, f);
  st = parallel_write(f, m.n_quads, 1u << 16, [&](uint64_t b, uint64_t e, std::string& out) {
    out.resize((size_t)(e - b) * 100);
    char* o = &out[0];
    for (uint64_t q = b; q < e; ++q)
      for (int t = 0; t < 2; ++t) {
        uint64_t tri[3];
        tri_of_quad(m.quads + 4 * q, t, tri);
        const float* p[3] = {m.positions + 3 * tri[0], m.positions + 3 * tri[1], m.positions + 3 * tri[2]};
        float n[3];
        tri_normal(p[0], p[1], p[2], n);
        const float len = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
        if (len > 0.0f) { n[0] /= len; n[1] /= len; n[2] /= len; } else { n[0] = n[1] = n[2] = 0.0f; }
        memcpy(o, n, 12); o += 12;
        for (int k = 0; k < 3; ++k) { memcpy(o, p[k], 12); o += 12; }
        o[0] = o[1] = 0; o += 2;
      }
  });
  if (fclose(f) != 0 && st == S2M_OK) st = fail(S2M_ERR_IO, std::string("close failed for ") + path);
  return st;
}
