/* s2m_sdf3d_lib.h -- the reference's built-in SDF libraries as CUDA device functions.
 *
 * Hand-written replacements (operation order preserved) for the three WGSL modules that
 * `use sdf3d::*;`, `use sdf3d::primitives;`, `use sdf3d::normal;`, `use sdf::op;`, `use sdf::*;`
 * pull into a .sdf3d file (/root/reference/src/shader.rs:12-20, :50-62):
 *   /root/reference/src/sdf3d_primitives.wgsl:7-36   box / cylinder / capsule / sphere / torus
 *   /root/reference/src/sdf_op.wgsl:7-23             smooth union / intersection / subtraction
 *   /root/reference/src/sdf3d_normal.wgsl:4-10       tetrahedral 4-tap gradient (templated on the SDF)
 * The front-end does not re-emit functions that come from those modules; JIT-compiled SDF code
 * links against these instead.
 */
#ifndef S2M_SDF3D_LIB_H_
#define S2M_SDF3D_LIB_H_
#include "s2m_vec.h"

namespace s2m {

/* sdf3d_primitives.wgsl:7-11 */
S2M_HD float sdf3d_box(vec3 p, vec3 b) {
  vec3 q = f_abs(p) - 0.5f * b;
  return f_length(f_max(q, mk3(0.0f, 0.0f, 0.0f))) + s2m_min(s2m_max(q.x, s2m_max(q.y, q.z)), 0.0f);
}
/* :13-17 */
S2M_HD float sdf3d_cylinder(vec3 p, float h, float r) {
  vec2 d = f_abs(mk2(f_length(mk2(p.x, p.z)), p.y)) - mk2(r, h);
  return s2m_min(s2m_max(d.x, d.y), 0.0f) + f_length(f_max(d, mk2(0.0f, 0.0f)));
}
/* :19-25 */
S2M_HD float sdf3d_capsule(vec3 p, vec3 a, vec3 b, float r) {
  vec3 pa = p - a;
  vec3 ba = b - a;
  float h = s2m_clamp(f_dot(pa, ba) / f_dot(ba, ba), 0.0f, 1.0f);
  return f_length(pa - ba * h) - r;
}
/* :27-30 */
S2M_HD float sdf3d_sphere(vec3 p, float s) { return f_length(p) - s; }
/* :32-36 */
S2M_HD float sdf3d_torus(vec3 p, vec2 t) {
  vec2 q = mk2(f_length(mk2(p.x, p.z)) - t.x, p.y);
  return f_length(q) - t.y;
}

/* sdf_op.wgsl:7-11 */
S2M_HD float sdf_op_smooth_union(float d1, float d2, float k) {
  float h = s2m_clamp(0.5f + 0.5f * (d2 - d1) / k, 0.0f, 1.0f);
  return s2m_mix(d2, d1, h) - k * h * (1.0f - h);
}
/* :13-17 */
S2M_HD float sdf_op_smooth_intersection(float d1, float d2, float k) {
  float h = s2m_clamp(0.5f - 0.5f * (d2 - d1) / k, 0.0f, 1.0f);
  return s2m_mix(d2, d1, h) + k * h * (1.0f - h);
}
/* :19-23 */
S2M_HD float sdf_op_smooth_subtraction(float d1, float d2, float k) {
  float h = s2m_clamp(0.5f - 0.5f * (d2 + d1) / k, 0.0f, 1.0f);
  return s2m_mix(d2, -d1, h) + k * h * (1.0f - h);
}

}  // namespace s2m
#endif
