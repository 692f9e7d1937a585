// common.h -- internal declarations shared by engine.cpp, shader_api.cpp and the front-end.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/sdf2mesh_b200.h"

namespace s2m_internal {

// thread-local error slot behind s2m_last_error()
void set_error(const std::string& msg);
int fail(int status, const std::string& msg);

// sources embedded at build time (tools/embed.py -> embedded_sources.cpp)
extern const char* const kSrcMathH;
extern const char* const kSrcVecH;
extern const char* const kSrcSdfLibH;
extern const char* const kSrcPvecH;
extern const char* const kSrcScanCuh;
extern const char* const kSrcKernelsJit;

}  // namespace s2m_internal

// Sdf3DShader (/root/reference/src/shader.rs:35-40)
struct s2m_shader {
  std::string source;    // the assembled source string
  int kind = S2M_SRC_SDF3D;
  std::string sdf_name;  // GLSL: name of the `float f(vec3)` function to wrap as sdf3d
  std::string glsl;      // GLSL input text (kind == S2M_SRC_GLSL_FRAGMENT)
  std::string log;       // what the reference would have logged
  // names of functions that came from a built-in module (`use sdf3d::*;` ...): they are linked
  // from s2m_sdf3d_lib.h / kernels_jit.cuh instead of being re-emitted
  std::vector<std::string> builtin_functions;
};

namespace s2m_frontend {
// Lower an assembled shader to CUDA C++ (the body of namespace s2m_user).  Returns s2m_status.
int lower_to_cuda(const s2m_shader& sh, std::string* cuda, std::string* err, std::string* packed);
}
