// frontend.cpp -- front-end entry points.
#include "frontend.h"

#include <cstdlib>

#include "parse.h"

namespace s2m_frontend {

int lower_to_cuda(const s2m_shader& sh, std::string* cuda, std::string* err, std::string* packed) {
  try {
    if (packed) packed->clear();
    if (sh.kind == S2M_SRC_CUDA) { *cuda = sh.source; return S2M_OK; }
    Module m;
    parse_wgsl(sh.source, sh.builtin_functions, &m);
    if (!getenv("S2M_NO_IR_OPT")) optimize_module(m);
    *cuda = emit_cuda(m);
    if (packed) *packed = emit_cuda_packed(m);
    return S2M_OK;
  } catch (const FrontendError& e) {
    *err = e.what();
    return e.status;
  } catch (const std::exception& e) {
    *err = std::string("front-end: ") + e.what();
    return S2M_ERR_SHADER;
  }
}

int glsl_to_wgsl(const std::string& glsl, std::string* wgsl, std::string* err) {
  try {
    Module m;
    parse_glsl(glsl, &m);
    check_eager_conditionals(m);
    make_names_wgsl_safe(m);
    *wgsl = emit_wgsl(m);
    return S2M_OK;
  } catch (const FrontendError& e) {
    *err = e.what();
    return e.status;
  } catch (const std::exception& e) {
    *err = std::string("front-end: ") + e.what();
    return S2M_ERR_SHADER;
  }
}

}  // namespace s2m_frontend
