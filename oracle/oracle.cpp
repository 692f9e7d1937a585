// oracle/oracle.cpp -- CPU ORACLE.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// build, load or call this.  Nothing under sdf2mesh_b200/ links or imports it; the product path
// fails loudly without its CUDA library.
//
// What it is: a literal restatement, in scalar strict-f32 C++ (g++ -O2 -ffp-contract=off -mfma),
// of the reference's SDF -> dual-contoured quads path exactly as the reference executes it:
//   /root/reference/src/bin/sdf2mesh/dualcontour.wgsl   (whole file: cell_bounds :22-27, cell_new
//       :29-43 = 8 SDF evaluations per cell, sign nibble :57-69, _cell_adapt/_cell_change :72-83,
//       cell_fetch_interpolated_pos :86-131, entry main :161-180)
//   /root/reference/src/sdf3d_normal.wgsl:4-10
//   /root/reference/src/bin/sdf2mesh/main.rs:139-175 (grid/bounds/eps), :298-356 (slice loop,
//       pixel scan order, `p.3 > 0.0` test, one-slice readback lag -- SURVEY.md F3)
//   /root/reference/src/mesh.rs:224-226 (key), :267-331 (quads by binary search)
//   /root/reference/src/lib.rs:115-118 (Bounds3D::cube), :187-211 (Quad swap / triangles / validity)
//   /root/reference/src/mesh.rs:8-48, :50-141, :167-210 (ASCII STL / PLY text)
//
// PARITY UNPINNED: the reference ships no golden vectors, known-answer tests or fixtures for this
// path, and it cannot be built or run in this image (no Rust, no Vulkan /
// lavapipe).  The known-answer set under tests/golden/ is therefore produced by THIS restatement
// (tests/golden/make_golden.py), not by the reference binary.  The float semantics the reference
// leaves to its driver compiler are pinned as listed in sdf_examples.h.
//
// Reference-cost mode is the only mode: 8 evaluations per cell + 4 per vertex, slice by slice,
// R*R pixel scan per slice, binary-search quad assembly.  std::thread workers over y rows inside a
// slice (this image has no libgomp).
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <atomic>
#include <chrono>
#include <thread>
#include "sdf_examples.h"

using osdf::V3;

namespace {

struct Item {            // mesh.rs:213-217 VertexListItem
  uint16_t x, y, z;      // z is the LABEL the reference stores (true slice + 1 in faithful mode)
  uint8_t nibble;        // sign_changes: bit0 s100, bit1 s010, bit2 s001, bit3 s000 (main.rs:338-339)
  float pos[3], nrm[3];
  uint64_t key() const { return (uint64_t)x | ((uint64_t)y << 16) | ((uint64_t)z << 32); }  // mesh.rs:224-226
};

struct Mesh {
  std::vector<Item> items;
  std::vector<uint64_t> quads;  // 4 per quad, after swap, emission order (valid quads only)
  uint64_t n_invalid = 0;
  std::vector<uint64_t> invalid;  // 6 per invalid quad: key, edge (0 X, 1 Y, 2 Z), q0..q3 after swap (MISSING = ~0)
  double seconds_cells = 0, seconds_quads = 0;
};

struct Grid {
  uint32_t res[3];
  float bmin[3], bmax[3], eps;
};

const uint32_t FLAG_ALL_SLICES = 1u;  // scan every slice, label = true z (no F3 lag)
const uint32_t FLAG_CONSISTENT_CORNERS = 64u;  // NOT the reference: cell max = next cell's min (S2M_MESH_CONSISTENT_CORNERS)
const uint64_t MISSING = ~0ull;

// dualcontour.wgsl:22-27
inline void cell_bounds(const Grid& g, int x, int y, int z, float cmin[3], float cmax[3], bool consistent = false) {
  const int pos[3] = {x, y, z};
  for (int a = 0; a < 3; ++a) {
    float v = (float)(g.res[a] - 1u);
    float size = (g.bmax[a] - g.bmin[a]) / v;
    cmin[a] = g.bmin[a] + size * (float)pos[a];
    cmax[a] = consistent ? g.bmin[a] + size * (float)(pos[a] + 1) : cmin[a] + size;
  }
}

// dualcontour.wgsl:72-83
inline float cell_adapt(float v0, float v1) { return (0.0f - v0) / (v1 - v0); }
inline void cell_change(float a, float b, float x, float y, float z, float out[3]) {
  if ((a > 0.0f) != (b > 0.0f)) { out[0] = x; out[1] = y; out[2] = z; }
  else { out[0] = 0.0f; out[1] = 0.0f; out[2] = 0.0f; }
}

// One invocation of the compute shader entry point (dualcontour.wgsl:161-180).
// Returns true iff the host would accept the pixel (main.rs:331: p.3 > 0.0).
inline bool run_cell(int sdf, const Grid& g, int x, int y, int z, Item& it, bool consistent = false) {
  float cmin[3], cmax[3];
  cell_bounds(g, x, y, z, cmin, cmax, consistent);
  // cell_new :29-43
  float d[8];
  for (int c = 0; c < 8; ++c) {
    V3 p = {(c & 1) ? cmax[0] : cmin[0], (c & 2) ? cmax[1] : cmin[1], (c & 4) ? cmax[2] : cmin[2]};
    d[c] = osdf::eval(sdf, p);
  }
  const float c000 = d[0], c100 = d[1], c010 = d[2], c110 = d[3], c001 = d[4], c101 = d[5], c011 = d[6], c111 = d[7];
  // cell_fetch_interpolated_pos :86-131
  float ch[12][3];
  cell_change(c000, c001, 0.0f, 0.0f, cell_adapt(c000, c001), ch[0]);
  cell_change(c010, c011, 0.0f, 1.0f, cell_adapt(c010, c011), ch[1]);
  cell_change(c100, c101, 1.0f, 0.0f, cell_adapt(c100, c101), ch[2]);
  cell_change(c110, c111, 1.0f, 1.0f, cell_adapt(c110, c111), ch[3]);
  cell_change(c000, c010, 0.0f, cell_adapt(c000, c010), 0.0f, ch[4]);
  cell_change(c001, c011, 0.0f, cell_adapt(c001, c011), 1.0f, ch[5]);
  cell_change(c100, c110, 1.0f, cell_adapt(c100, c110), 0.0f, ch[6]);
  cell_change(c101, c111, 1.0f, cell_adapt(c101, c111), 1.0f, ch[7]);
  cell_change(c000, c100, cell_adapt(c000, c100), 0.0f, 0.0f, ch[8]);
  cell_change(c001, c101, cell_adapt(c001, c101), 0.0f, 1.0f, ch[9]);
  cell_change(c010, c110, cell_adapt(c010, c110), 1.0f, 0.0f, ch[10]);
  cell_change(c011, c111, cell_adapt(c011, c111), 1.0f, 1.0f, ch[11]);
  float avg[3] = {0.0f, 0.0f, 0.0f};
  float count = 0.0f;
  for (int i = 0; i < 12; ++i) {
    if (ch[i][0] > 0.0f || ch[i][1] > 0.0f || ch[i][2] > 0.0f) {
      avg[0] += ch[i][0]; avg[1] += ch[i][1]; avg[2] += ch[i][2];
      count += 1.0f;
    }
  }
  if (count <= 1.0f) return false;  // vec4(-1): shader stores zeros, host test p.w > 0 fails
  float pos[3];
  for (int a = 0; a < 3; ++a) pos[a] = cmin[a] + (cmax[a] - cmin[a]) * avg[a] / count;  // :130
  // sdf3d_normal.wgsl:4-10, then normalize (dualcontour.wgsl:171)
  const float v[4][3] = {{1.0f, -1.0f, -1.0f}, {-1.0f, -1.0f, 1.0f}, {-1.0f, 1.0f, -1.0f}, {1.0f, 1.0f, 1.0f}};
  float n[3] = {0, 0, 0};
  for (int k = 0; k < 4; ++k) {
    V3 q = {pos[0] + v[k][0] * g.eps, pos[1] + v[k][1] * g.eps, pos[2] + v[k][2] * g.eps};
    float f = osdf::eval(sdf, q);
    if (k == 0) { n[0] = v[k][0] * f; n[1] = v[k][1] * f; n[2] = v[k][2] * f; }
    else { n[0] = n[0] + v[k][0] * f; n[1] = n[1] + v[k][1] * f; n[2] = n[2] + v[k][2] * f; }
  }
  float len = sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
  for (int a = 0; a < 3; ++a) { it.pos[a] = pos[a]; it.nrm[a] = n[a] / len; }
  // cell_sign_changes_f32 :57-69
  float s = 0.0f;
  if (c100 > 0.0f) s += 1.0f;
  if (c010 > 0.0f) s += 2.0f;
  if (c001 > 0.0f) s += 4.0f;
  if (c000 > 0.0f) s += 8.0f;
  it.nibble = (uint8_t)(uint32_t)s;
  it.x = (uint16_t)x; it.y = (uint16_t)y;
  return true;
}

// mesh.rs:327-331
inline uint64_t vertex_index(const std::vector<Item>& v, uint16_t x, uint16_t y, uint16_t z) {
  uint64_t key = (uint64_t)x | ((uint64_t)y << 16) | ((uint64_t)z << 32);
  size_t lo = 0, hi = v.size();
  while (lo < hi) {
    size_t mid = (lo + hi) / 2;
    uint64_t k = v[mid].key();
    if (k == key) return mid;
    if (k < key) lo = mid + 1; else hi = mid;
  }
  return MISSING;
}

// mesh.rs:267-324 + lib.rs:190-211
void assemble_quads(Mesh& m) {
  uint64_t cur_key = 0, cur_edge = 0;
  auto push = [&](uint64_t q0, uint64_t q1, uint64_t q2, uint64_t q3, bool swap) {
    uint64_t q[4] = {q0, q1, q2, q3};
    if (swap) { std::swap(q[0], q[3]); std::swap(q[1], q[2]); }
    if (q[0] != MISSING && q[1] != MISSING && q[2] != MISSING && q[3] != MISSING) m.quads.insert(m.quads.end(), q, q + 4);
    else {
      m.n_invalid++;  // mesh.rs:276 log::warn!("Invalid quad: {:?}. Mesh will not be water-tight!", quad)
      const uint64_t rec[6] = {cur_key, cur_edge, q[0], q[1], q[2], q[3]};
      m.invalid.insert(m.invalid.end(), rec, rec + 6);
    }
  };
  const std::vector<Item>& v = m.items;
  for (size_t i = 0; i < v.size(); ++i) {
    const Item& it = v[i];
    bool s100 = it.nibble & 1, s010 = it.nibble & 2, s001 = it.nibble & 4, s000 = it.nibble & 8;
    uint16_t x = it.x, y = it.y, z = it.z;
    cur_key = it.key();
    cur_edge = 0;
    if (s100 != s000 && y > 0 && z > 0)
      push(vertex_index(v, x, y - 1, z - 1), vertex_index(v, x, y, z - 1), vertex_index(v, x, y, z), vertex_index(v, x, y - 1, z), s100);
    cur_edge = 1;
    if (s010 != s000 && x > 0 && z > 0)
      push(vertex_index(v, x - 1, y, z - 1), vertex_index(v, x, y, z - 1), vertex_index(v, x, y, z), vertex_index(v, x - 1, y, z), !s010);
    cur_edge = 2;
    if (s001 != s000 && x > 0 && y > 0)
      push(vertex_index(v, x - 1, y - 1, z), vertex_index(v, x, y - 1, z), vertex_index(v, x, y, z), vertex_index(v, x - 1, y, z), s001);
  }
}

double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// Rust `{}` for f32: shortest digits that round-trip, plain decimal, no exponent.
std::string rust_f32(float f) {
  if (f != f) return "NaN";
  if (f == s2m_inf()) return "inf";
  if (f == -s2m_inf()) return "-inf";
  char buf[64];
  int prec = 1;
  for (; prec <= 9; ++prec) {
    snprintf(buf, sizeof buf, "%.*e", prec - 1, (double)f);
    if (strtof(buf, nullptr) == f) break;
  }
  // buf = [-]d.ddddde[+-]XX
  std::string s(buf);
  bool neg = s[0] == '-';
  if (neg) s = s.substr(1);
  size_t epos = s.find('e');
  int ex = atoi(s.c_str() + epos + 1);
  std::string digits;
  for (size_t i = 0; i < epos; ++i) if (s[i] != '.') digits += s[i];
  while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
  std::string out;
  if (digits == "0") out = "0";
  else if (ex < 0) out = "0." + std::string((size_t)(-ex - 1), '0') + digits;
  else if ((size_t)ex + 1 >= digits.size()) out = digits + std::string((size_t)ex + 1 - digits.size(), '0');
  else out = digits.substr(0, (size_t)ex + 1) + "." + digits.substr((size_t)ex + 1);
  return neg ? "-" + out : out;
}

}  // namespace

extern "C" {

int oracle_num_threads() {
  unsigned n = std::thread::hardware_concurrency();
  return n ? (int)n : 1;
}

void oracle_set_plugin(void* fn) { osdf::g_plugin = reinterpret_cast<osdf::PluginFn>(fn); }

void oracle_eval(int sdf, const float* pts, float* out, uint64_t n) {
  for (uint64_t i = 0; i < n; ++i) out[i] = osdf::eval(sdf, V3{pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]});
}

// corner coordinates along one axis as the shader computes them for cell j:
// a[j] = cell j's min, b[j] = cell j's max (= a[j] + size, NOT a[j+1] in general: SURVEY.md F4)
void oracle_axis_coords(uint32_t res, float bmin, float bmax, float* a, float* b) {
  float size = (bmax - bmin) / (float)(res - 1u);
  for (uint32_t j = 0; j < res; ++j) { a[j] = bmin + size * (float)j; b[j] = a[j] + size; }
}

// Run the path over true cell slices [z_begin, z_end) (0,0 = the reference's full range).
// threads <= 0: all OpenMP threads.
void* oracle_mesh_run(int sdf, const uint32_t res[3], const float bmin[3], const float bmax[3], float eps,
                      uint32_t flags, uint32_t z_begin, uint32_t z_end, int threads) {
  Grid g;
  for (int a = 0; a < 3; ++a) { g.res[a] = res[a]; g.bmin[a] = bmin[a]; g.bmax[a] = bmax[a]; }
  g.eps = eps;
  const bool all = flags & FLAG_ALL_SLICES;
  // faithful: the scan at loop iteration z sees slice z-1 (main.rs:321-325 vs :355); slice res-1 is never read
  uint32_t zlast = all ? g.res[2] : g.res[2] - 1;
  if (z_begin == 0 && z_end == 0) z_end = zlast;
  z_end = std::min(z_end, zlast);
  if (threads <= 0) threads = oracle_num_threads();
  Mesh* m = new Mesh();
  double t0 = now_s();
  const int ny = (int)g.res[1], nx = (int)g.res[0];
  std::vector<std::vector<Item>> rows((size_t)ny);
  for (uint32_t z = z_begin; z < z_end; ++z) {
    std::atomic<int> next{0};
    auto work = [&]() {
      for (;;) {
        int y0 = next.fetch_add(4);
        if (y0 >= ny) break;
        for (int y = y0; y < std::min(ny, y0 + 4); ++y) {
          rows[(size_t)y].clear();
          for (int x = 0; x < nx; ++x) {
            Item it;
            if (run_cell(sdf, g, x, y, (int)z, it, (flags & FLAG_CONSISTENT_CORNERS) != 0)) {
              it.z = (uint16_t)(all ? z : z + 1);
              rows[(size_t)y].push_back(it);
            }
          }
        }
      }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(work);
    work();
    for (auto& th : pool) th.join();
    for (int y = 0; y < ny; ++y) m->items.insert(m->items.end(), rows[(size_t)y].begin(), rows[(size_t)y].end());
  }
  double t1 = now_s();
  assemble_quads(*m);
  double t2 = now_s();
  m->seconds_cells = t1 - t0;
  m->seconds_quads = t2 - t1;
  return m;
}

void oracle_mesh_counts(void* h, uint64_t* nv, uint64_t* nq, uint64_t* ninv, double* sec_cells, double* sec_quads) {
  Mesh* m = (Mesh*)h;
  *nv = m->items.size(); *nq = m->quads.size() / 4; *ninv = m->n_invalid;
  *sec_cells = m->seconds_cells; *sec_quads = m->seconds_quads;
}

void oracle_mesh_copy(void* h, float* pos, float* nrm, uint64_t* keys, uint8_t* nibbles, uint64_t* quads) {
  Mesh* m = (Mesh*)h;
  for (size_t i = 0; i < m->items.size(); ++i) {
    const Item& it = m->items[i];
    for (int a = 0; a < 3; ++a) { pos[3 * i + a] = it.pos[a]; nrm[3 * i + a] = it.nrm[a]; }
    keys[i] = it.key();
    nibbles[i] = it.nibble;
  }
  if (!m->quads.empty()) memcpy(quads, m->quads.data(), m->quads.size() * 8);
}

uint64_t oracle_mesh_invalid(void* h, uint64_t* out) {  // out may be NULL: returns the record count
  Mesh* m = (Mesh*)h;
  if (out && !m->invalid.empty()) memcpy(out, m->invalid.data(), m->invalid.size() * 8);
  return m->invalid.size() / 6;
}

void oracle_mesh_free(void* h) { delete (Mesh*)h; }

// mesh.rs:12-48 + :155-180 + lib.rs:181-185.  Triangles (q2,q1,q0),(q0,q3,q2) per quad (lib.rs:199-204).
int oracle_write_stl(void* h, const char* path) {
  Mesh* m = (Mesh*)h;
  FILE* f = fopen(path, "wb");
  if (!f) return 1;
  fputs("solid\n", f);
  for (size_t q = 0; q + 3 < m->quads.size(); q += 4) {
    const uint64_t* Q = &m->quads[q];
    const uint64_t tri[2][3] = {{Q[2], Q[1], Q[0]}, {Q[0], Q[3], Q[2]}};
    for (int t = 0; t < 2; ++t) {
      const float* p0 = m->items[tri[t][0]].pos; const float* p1 = m->items[tri[t][1]].pos; const float* p2 = m->items[tri[t][2]].pos;
      float a[3] = {p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2]};
      float b[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]};
      float n[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};  // euclid cross
      fprintf(f, "facet normal %s %s %s\n", rust_f32(n[0]).c_str(), rust_f32(n[1]).c_str(), rust_f32(n[2]).c_str());
      fputs("\touter loop\n", f);
      for (const float* p : {p0, p1, p2})
        fprintf(f, "\t\tvertex %s %s %s\n", rust_f32(p[0]).c_str(), rust_f32(p[1]).c_str(), rust_f32(p[2]).c_str());
      fputs("\tendloop\nendfacet\n", f);
    }
  }
  fputs("endsolid\n", f);
  fclose(f);
  return 0;
}

// mesh.rs:55-99, :130-141, :198-210
int oracle_write_ply(void* h, const char* path) {
  Mesh* m = (Mesh*)h;
  FILE* f = fopen(path, "wb");
  if (!f) return 1;
  fprintf(f, "ply\nformat ascii 1.0\ncomment written by rust-sdf\n");
  fprintf(f, "element vertex %zu\nproperty float x\nproperty float y\nproperty float z\nproperty float nx\nproperty float ny\nproperty float nz\n", m->items.size());
  fprintf(f, "element face %zu\nproperty list uchar int vertex_index\nend_header\n", m->quads.size() / 2);
  for (const Item& it : m->items)
    fprintf(f, "%s %s %s %s %s %s\n", rust_f32(it.pos[0]).c_str(), rust_f32(it.pos[1]).c_str(), rust_f32(it.pos[2]).c_str(),
            rust_f32(it.nrm[0]).c_str(), rust_f32(it.nrm[1]).c_str(), rust_f32(it.nrm[2]).c_str());
  for (size_t q = 0; q + 3 < m->quads.size(); q += 4) {
    const uint64_t* Q = &m->quads[q];
    fprintf(f, "3 %u %u %u\n3 %u %u %u\n", (unsigned)Q[2], (unsigned)Q[1], (unsigned)Q[0], (unsigned)Q[0], (unsigned)Q[3], (unsigned)Q[2]);
  }
  fclose(f);
  return 0;
}

const char* oracle_rust_f32(float f) {
  static thread_local std::string s;
  s = rust_f32(f);
  return s.c_str();
}

}  // extern "C"
