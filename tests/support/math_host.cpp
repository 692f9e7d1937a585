// Host build of the pinned math (tests only): array maps over s2m_math.h functions.
#include "math_dispatch.h"
#include <cstddef>
extern "C" void s2m_host_map1(int fn, const float* x, float* y, size_t n) {
  for (size_t i = 0; i < n; ++i) y[i] = s2m_dispatch1(fn, x[i]);
}
extern "C" void s2m_host_map2(int fn, const float* a, const float* b, float* y, size_t n) {
  for (size_t i = 0; i < n; ++i) y[i] = s2m_dispatch2(fn, a[i], b[i]);
}
