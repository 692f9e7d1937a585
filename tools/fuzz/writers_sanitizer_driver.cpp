// The mesh writers (csrc/writers.cpp: worker threads formatting chunks into a ring of buffers, one thread writing
// them in order) under ThreadSanitizer and AddressSanitizer, on a synthetic mesh; tools/fuzz/run_writers_sanitizers.sh
// runs it with several thread counts and chunk sizes and compares the files.  Stubs what writers.cpp uses from
// engine.cpp.  Not product code.
#include "common.h"
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
namespace s2m_internal {
static thread_local std::string g_err;
void set_error(const std::string& m) { g_err = m; }
int fail(int st, const std::string& m) { g_err = m; return st; }
}
extern "C" const char* s2m_last_error(void) { return s2m_internal::g_err.c_str(); }
extern "C" int s2m_result_get(const s2m_result*, s2m_result_info*) { return 1; }
int main(int argc, char** argv) {
  const uint64_t nv = 20000, nq = 19000;
  std::vector<float> pos(3 * nv), nrm(3 * nv);
  for (uint64_t i = 0; i < 3 * nv; ++i) { pos[i] = std::sin(0.37f * i) * 3.0f; nrm[i] = std::cos(0.11f * i); }
  std::vector<uint64_t> q(4 * nq);
  for (uint64_t i = 0; i < 4 * nq; ++i) q[i] = (i * 2654435761ull) % nv;
  s2m_result_info info{};
  info.n_vertices = nv; info.n_quads = nq; info.positions = pos.data(); info.normals = nrm.data(); info.quads = q.data();
  for (const char* path : {"a.stl", "a.ply"}) {
    int st = s2m_write_mesh_arrays(&info, 1, path, 0);
    if (st) { printf("error %d %s\n", st, s2m_last_error()); return 1; }
  }
  int st = s2m_write_mesh_arrays(&info, 1, "b.stl", 1);
  printf("done %d\n", st);
  return st;
}
