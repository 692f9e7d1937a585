#include "parse.h"
namespace s2m_frontend {
void parse_glsl(const std::string&, Module*) { throw FrontendError(11, "GLSL front-end not built yet"); }
}
