// more idioms: arrays of structs, out params, globals written by functions, nested functions, vec4 swizzles rgba/stpq,
// integer for with multiple vars, ternary chains, early returns, compound ops on swizzles, const arrays of vec2, mat4, etc.
#define saturate(x) clamp(x,0.0,1.0)
#define S(a,b,t) smoothstep(a,b,t)
struct Prim { vec3 c; float r; int mat; };
const Prim prims[3] = Prim[3](Prim(vec3(0.), 1.0, 0), Prim(vec3(1.,0.,0.), .5, 1), Prim(vec3(0.,1.2,0.), .25, 2));
const vec2 offs[4] = vec2[](vec2(1,0), vec2(0,1), vec2(-1,0), vec2(0,-1));
float gMin = 1e9;
int gMat = -1;
void track(float d, int m) { if (d < gMin) { gMin = d; gMat = m; } }
vec4 quatMul(vec4 a, vec4 b) { return vec4(a.w*b.xyz + b.w*a.xyz + cross(a.xyz, b.xyz), a.w*b.w - dot(a.xyz, b.xyz)); }
vec3 rotate(vec3 v, vec4 q) { vec4 t = quatMul(quatMul(q, vec4(v, 0.)), vec4(-q.xyz, q.w)); return t.rgb; }
mat4 translate(vec3 t) { mat4 m = mat4(1.0); m[3] = vec4(t, 1.0); return m; }
float capsule(vec3 p, vec3 a, vec3 b, float r) { vec3 pa = p - a, ba = b - a; float h = saturate(dot(pa,ba)/dot(ba,ba)); return length(pa - ba*h) - r; }
void sincosf(float a, out float s, out float c) { s = sin(a); c = cos(a); }
float ellipsoid(in vec3 p, in vec3 r) { float k0 = length(p/r); float k1 = length(p/(r*r)); return k0*(k0-1.0)/k1; }
float sdf(vec3 p) {
    gMin = 1e9; gMat = -1;
    vec4 q = normalize(vec4(0.1, 0.2, 0.3, 1.0));
    vec3 pr = rotate(p, q);
    pr = (translate(vec3(0.1, 0., -0.2)) * vec4(pr, 1.0)).xyz;
    for (int i = 0, j = 2; i < 3; i++, j--) {
        Prim pm = prims[i];
        float d = length(pr - pm.c) - pm.r * (1.0 + 0.1*float(j));
        track(d, pm.mat);
    }
    float s, c;
    sincosf(p.y * 3.0, s, c);
    vec4 col = vec4(s, c, s*c, 1.0);
    col.rg *= 0.5; col.ba += col.rg;
    track(capsule(pr, vec3(-1,0,0), vec3(1,.5,0), .2) + 0.01*col.b, 3);
    for (int i = 0; i < 4; i++) { vec2 o = offs[i]; track(ellipsoid(pr - vec3(o*1.5, 0).xzy, vec3(.3,.2,.4)), 4 + i); }
    float e = gMat == 0 ? 0.01 : gMat == 1 ? 0.02 : gMat < 4 ? 0.03 : 0.04;
    if (gMin > 5.0) return gMin;
    return gMin - e * S(0., 1., abs(p.x));
}
void mainImage(out vec4 o, vec2 u) { o = vec4(sdf(vec3(u,0))); }
