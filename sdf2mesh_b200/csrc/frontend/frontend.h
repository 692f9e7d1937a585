// frontend.h -- entry points of the shader front-end (replaces naga: WGSL front-end used by
// wgpu's create_shader_module, /root/reference/src/shader.rs:220-225, and the GLSL front-end +
// WGSL writer used by convert_glsl_to_wgsl, /root/reference/src/shadertoy.rs:169-194).
#pragma once
#include <string>

#include "../common.h"

namespace s2m_frontend {
// assembled WGSL (or .sdf3d after directive expansion) -> body of `namespace s2m_user` in CUDA C++
// `packed` (optional): the same code over f32x2 pairs, body of `namespace s2m_user_p` (s2m_pvec.h); "" if not expressible
int lower_to_cuda(const s2m_shader& sh, std::string* cuda, std::string* err, std::string* packed = nullptr);
// GLSL fragment shader -> WGSL text (naga-shaped: user functions, `fn main_1()`, `@fragment fn main()`)
int glsl_to_wgsl(const std::string& glsl, std::string* wgsl, std::string* err);
}  // namespace s2m_frontend
