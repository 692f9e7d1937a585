#define time iTime
#define R(p,a) p=cos(a)*p+sin(a)*vec2(p.y,-p.x)
#define MAT_COUNT 4
#if MAT_COUNT > 2 && defined(time)
#define USE_HASH 1
#else
#define USE_HASH 0
#endif
#define LONG_MACRO(a, b) \
    ((a) * (b) + \
     (a))
float sdf(vec3 p);   // prototype
const float weights[] = float[](0.5, 0.25, 0.125, 0.0625);
uvec3 pcg3d(uvec3 v) {
    v = v * 1664525u + 1013904223u;
    v.x += v.y*v.z; v.y += v.z*v.x; v.z += v.x*v.y;
    v ^= v >> 16u;
    v.x += v.y*v.z; v.y += v.z*v.x; v.z += v.x*v.y;
    return v;
}
vec3 hash33(vec3 p) { return vec3(pcg3d(floatBitsToUint(p))) * (1.0/float(0xffffffffu)); }
float fbm(vec3 p) {
    float a = 0.0;
    for (int i = 0; i < weights.length(); i++) { a += weights[i] * length(hash33(floor(p))); p = p * 2.0 + 1.0; }
    return a;
}
float vmax(vec3 v) { return max(max(v.x, v.y), v.z); }
float sgn(float x) { return (x<0.)?-1.:1.; }
float pModPolar(inout vec2 p, float repetitions) {
    float angle = 2.*3.14159265/repetitions;
    float a = atan(p.y, p.x) + angle/2.;
    float r = length(p);
    float c = floor(a/angle);
    a = mod(a,angle) - angle/2.;
    p = vec2(cos(a), sin(a))*r;
    if (abs(c) >= (repetitions/2.)) c = abs(c);
    return c;
}
float shape(vec3 p, int kind) {
    switch (kind) {
        case 0: return length(p) - 1.0;
        case 1: { vec3 d = abs(p) - vec3(0.7); return vmax(d); }
        case 2:
        case 3: return length(p.xy) - 0.5;
        default: break;
    }
    return 1e10;
}
float sdf(vec3 p) {
    vec3 q = p;
    R(q.xy, 0.5);
    float cell = pModPolar(q.xz, 6.0);
    float d = 1e10;
    float arr[MAT_COUNT];
    for (int i = 0; i < MAT_COUNT; i++) arr[i] = shape(q - vec3(1.5, 0., 0.), i);
    int k = 0;
    do { d = min(d, arr[k]); k++; } while (k < MAT_COUNT);
    for (float f = 0.; f < 1.; f += .25) d = min(d, length(p - vec3(f, 2.0*f, 0.)) - .1);
    mat3 m = mat3(1.0);
    m[1] = vec3(0., 2., 0.);
    m[2][0] = 0.5;
    vec3 w = m * p;
    w[k % 3] += 0.25;
    bvec3 pos = greaterThan(w, vec3(0.));
    d += all(pos) ? 0.01 : (any(pos) ? 0.02 : 0.03);
    ivec2 ij = ivec2(floor(p.xy));
    d += float((ij.x ^ ij.y) & 1) * 0.001;
#if USE_HASH
    d += 0.01 * fbm(p * 3.0) + LONG_MACRO(0.001, cell);
#endif
    int n = 0;
    while (n < 3) { if (d > float(n)) { n += 2; continue; } n++; }
    return d - sgn(p.y) * 0.01 * float(n);
}
void mainImage(out vec4 o, in vec2 u) { o = vec4(sdf(vec3(u, time))); }
