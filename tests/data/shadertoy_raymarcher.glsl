// A ShaderToy-style raymarcher written for this test suite (not taken from the reference): the
// constructs such shaders use -- #define constants, structs with ?: on them, mat2 rotation,
// smooth min, domain repetition with mod, value noise, swizzle stores, mainImage with the
// ShaderToy uniforms.
#define MAX_STEPS 100
#define PI 3.14159265
#define SURF_DIST .001
const float MAX_DIST = 100.;

struct Hit { float d; int id; };

mat2 Rot(float a) { float s = sin(a), c = cos(a); return mat2(c, -s, s, c); }
float smin(float a, float b, float k) { float h = clamp(0.5 + 0.5*(b-a)/k, 0., 1.); return mix(b, a, h) - k*h*(1.0-h); }
float sdBox(vec3 p, vec3 s) { p = abs(p)-s; return length(max(p, 0.))+min(max(p.x, max(p.y, p.z)), 0.); }
float sdTorus(vec3 p, vec2 r) { float x = length(p.xz)-r.x; return length(vec2(x, p.y))-r.y; }
float sdCapsule(vec3 p, vec3 a, vec3 b, float r) { vec3 ab = b-a, ap = p-a; float t = clamp(dot(ab, ap)/dot(ab, ab), 0., 1.); return length(p - (a + t*ab)) - r; }
float hash(vec3 p) { p = fract(p * 0.3183099 + .1); p *= 17.0; return fract(p.x * p.y * p.z * (p.x + p.y + p.z)); }
float noise(in vec3 x) {
    vec3 i = floor(x); vec3 f = fract(x); f = f*f*(3.0-2.0*f);
    return mix(mix(mix(hash(i+vec3(0,0,0)), hash(i+vec3(1,0,0)), f.x), mix(hash(i+vec3(0,1,0)), hash(i+vec3(1,1,0)), f.x), f.y),
               mix(mix(hash(i+vec3(0,0,1)), hash(i+vec3(1,0,1)), f.x), mix(hash(i+vec3(0,1,1)), hash(i+vec3(1,1,1)), f.x), f.y), f.z);
}
Hit opU(Hit a, Hit b) { return (a.d < b.d) ? a : b; }
Hit mapHit(vec3 p) {
    vec3 q = p;
    q.xz *= Rot(iTime * .2 + 0.5);
    Hit h = Hit(sdBox(q, vec3(.5)), 1);
    vec3 rp = p; rp.x = mod(rp.x + 1.0, 2.0) - 1.0;
    h = opU(h, Hit(sdTorus(rp - vec3(0, .8, 0), vec2(.4, .1)), 2));
    h = opU(h, Hit(sdCapsule(p, vec3(-1, 0, 0), vec3(1, .5, .3), .15), 3));
    h.d = smin(h.d, length(p - vec3(0., -0.6, 0.)) - 0.4, 0.2);
    h.d += 0.03 * noise(p * 6.0);
    return h;
}
float map(vec3 p) { return mapHit(p).d; }
float RayMarch(vec3 ro, vec3 rd) {
    float dO = 0.;
    for (int i = 0; i < MAX_STEPS; i++) { vec3 p = ro + rd*dO; float dS = map(p); dO += dS; if (dO > MAX_DIST || abs(dS) < SURF_DIST) break; }
    return dO;
}
vec3 GetNormal(vec3 p) { vec2 e = vec2(.001, 0); vec3 n = map(p) - vec3(map(p-e.xyy), map(p-e.yxy), map(p-e.yyx)); return normalize(n); }
void mainImage(out vec4 fragColor, in vec2 fragCoord) {
    vec2 uv = (fragCoord - .5*iResolution.xy)/iResolution.y;
    vec3 ro = vec3(0, 3, -3); vec3 rd = normalize(vec3(uv, 1));
    float d = RayMarch(ro, rd);
    vec3 col = vec3(0);
    if (d < MAX_DIST) { vec3 p = ro + rd*d; vec3 n = GetNormal(p); col = n*.5+.5; }
    fragColor = vec4(pow(col, vec3(.4545)), 1.0);
}
