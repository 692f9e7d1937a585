#include "common.h"
extern "C" int s2m_result_write_mesh(const s2m_result*, const char*) { return s2m_internal::fail(S2M_ERR_UNSUPPORTED, "writers not built yet"); }
extern "C" int s2m_result_write_stl_binary(const s2m_result*, const char*) { return s2m_internal::fail(S2M_ERR_UNSUPPORTED, "writers not built yet"); }
