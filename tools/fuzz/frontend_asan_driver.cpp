// Front-end under AddressSanitizer + UndefinedBehaviorSanitizer: translates every file of a directory
// (*.glsl as GLSL fragment shaders, everything else as .sdf3d) -- tools/fuzz/run_frontend_asan.sh feeds it
// damaged copies of the examples and fixtures.  Stubs the two symbols of engine.cpp the front-end uses, so
// that no CUDA library is linked.  Not product code.
#include "common.h"
#include <cstdio>
#include <fstream>
#include <sstream>
#include <dirent.h>
namespace s2m_internal {
static thread_local std::string g_err;
void set_error(const std::string& m) { g_err = m; }
int fail(int st, const std::string& m) { g_err = m; return st; }
}
extern "C" const char* s2m_last_error(void) { return s2m_internal::g_err.c_str(); }
extern "C" const char* s2m_version(void) { return "asan"; }
extern "C" void s2m_free(void* p) { free(p); }
int main(int argc, char** argv) {
  DIR* d = opendir(argv[1]);
  int n = 0, ok = 0;
  while (dirent* e = readdir(d)) {
    std::string name = e->d_name;
    if (name.size() < 3) continue;
    std::ifstream f(std::string(argv[1]) + "/" + name, std::ios::binary);
    std::ostringstream ss; ss << f.rdbuf();
    std::string text = ss.str();
    const int kind = name.find(".glsl") != std::string::npos ? S2M_SRC_GLSL_FRAGMENT : S2M_SRC_SDF3D;
    s2m_shader* sh = nullptr;
    ++n;
    if (name.find(".json") != std::string::npos) {  // a ShaderToy API response
      if (s2m_shader_from_shadertoy_response(text.data(), text.size(), "sdf", &sh) == 0) { ++ok; s2m_shader_free(sh); }
      continue;
    }
    if (s2m_shader_from_source(text.data(), text.size(), kind, "sdf", nullptr, &sh) == 0) {
      char* c = nullptr;
      if (s2m_shader_lower_to_cuda(sh, &c) == 0) { ++ok; free(c); c = nullptr; if (s2m_shader_lower_to_cuda_packed(sh, &c) == 0) free(c); }
      s2m_shader_free(sh);
    }
  }
  closedir(d);
  printf("%d inputs, %d translated\n", n, ok);
}
