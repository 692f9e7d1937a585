// More ShaderToy idioms, written for this test suite: function-like macros, a global set per call,
// a constant array, an inout parameter fed with a swizzle (pModPolar(q.xz, 6.)), float loop
// counters, while(true) with break, folding fractal with swizzle swaps and mat2 rotation.
#define TAU 6.2831853
#define rot(a) mat2(cos(a), sin(a), -sin(a), cos(a))
#define sat(x) clamp(x, 0., 1.)
#define REP(p, c) (mod((p) + 0.5*(c), (c)) - 0.5*(c))
float gTime = 0.;
const int ITER = 5;
const vec3 OFFS[3] = vec3[3](vec3(1, 0, 0), vec3(0, 1, 0), vec3(0, 0, 1));

float sdRoundBox(vec3 p, vec3 b, float r) { vec3 q = abs(p) - b; return length(max(q, 0.0)) + min(max(q.x, max(q.y, q.z)), 0.0) - r; }
float sdHexPrism(vec3 p, vec2 h) {
  const vec3 k = vec3(-0.8660254, 0.5, 0.57735);
  p = abs(p);
  p.xy -= 2.0*min(dot(k.xy, p.xy), 0.0)*k.xy;
  vec2 d = vec2(length(p.xy - vec2(clamp(p.x, -k.z*h.x, k.z*h.x), h.x))*sign(p.y - h.x), p.z - h.y);
  return min(max(d.x, d.y), 0.0) + length(max(d, 0.0));
}
void pModPolar(inout vec2 p, float repetitions) {
  float angle = TAU/repetitions;
  float a = atan(p.y, p.x) + angle/2.;
  float r = length(p);
  a = mod(a, angle) - angle/2.;
  p = vec2(cos(a), sin(a))*r;
}
float fractal(vec3 p) {
  float s = 1.0;
  for (int i = 0; i < ITER; i++) {
    p = abs(p) - vec3(0.6, 0.4, 0.3) * s;
    if (p.x < p.y) p.xy = p.yx;
    if (p.x < p.z) p.xz = p.zx;
    p.yz *= rot(0.3 + float(i) * 0.1);
    s *= 0.6;
  }
  return sdRoundBox(p, vec3(0.1*s*4.0), 0.01);
}
float map(vec3 p) {
  gTime = iTime * 0.5;
  float d = 1e10;
  vec3 q = p;
  pModPolar(q.xz, 6.0);
  d = min(d, sdHexPrism(q - vec3(1.2, 0., 0.), vec2(0.2, 0.3)));
  for (float t = 0.; t < 1.; t += 0.34) d = min(d, length(p - vec3(0., t - 0.5, 0.)) - 0.15 * (1. - t));
  int k = 0;
  while (true) { d = min(d, length(p - OFFS[k] * 0.9) - 0.1); k++; if (k >= 3) break; }
  d = min(d, fractal(p * 1.3) / 1.3);
  vec3 r = REP(p, vec3(1.5));
  d = max(d, -(length(r) - 0.2 - 0.05 * sin(gTime)));
  return k % 2 == 1 ? d : d * 1.0;
}
void mainImage(out vec4 c, in vec2 f) { c = vec4(map(vec3(f / iResolution.xy, 0.)) > 0. ? 1. : 0.); }
