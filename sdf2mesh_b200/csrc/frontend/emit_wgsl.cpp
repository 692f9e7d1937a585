#include "parse.h"
namespace s2m_frontend {
std::string emit_wgsl(const Module&) { throw FrontendError(11, "WGSL writer not built yet"); }
}
