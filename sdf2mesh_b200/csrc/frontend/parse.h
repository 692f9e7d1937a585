// parse.h -- parser and writer entry points (internal to the front-end).
#pragma once
#include <string>
#include <vector>

#include "ir.h"

namespace s2m_frontend {
void parse_wgsl(const std::string& src, const std::vector<std::string>& builtin_fns, Module* out);
void parse_glsl(const std::string& src, Module* out);
int optimize_module(Module& m);  // returns the number of rewrites applied
void check_eager_conditionals(const Module& m);  // GLSL ?: left as select() must not hide a side effect (throws S2M_ERR_UNSUPPORTED)
std::string emit_cuda(const Module& m);
std::string emit_cuda_packed(const Module& m);  // "" when the module has no packed (f32x2) form
std::string emit_wgsl(const Module& m);
void make_names_wgsl_safe(Module& m);  // GLSL identifiers that WGSL reserves get a `_` suffix, as naga's namer does
}  // namespace s2m_frontend
