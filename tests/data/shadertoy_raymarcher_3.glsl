// typical ShaderToy raymarcher
#ifdef GL_ES
precision highp float;
#endif
#define AA 1
#define PI 3.14159265
#define ZERO (min(iFrame,0))
#define sat(x) clamp(x, 0., 1.)
const int MAX_STEPS = 128;
const float EPS = .001, FAR = 40.;
const vec3 cols[3] = vec3[3](vec3(1.,.5,.2), vec3(.2), vec3(0,1,0));
struct Hit { float d; int id; };

float dot2( in vec2 v ) { return dot(v,v); }
float dot2( in vec3 v ) { return dot(v,v); }
float ndot( in vec2 a, in vec2 b ) { return a.x*b.x - a.y*b.y; }
mat2 rot(float a) { float c = cos(a), s = sin(a); return mat2(c, -s, s, c); }
float hash(float n) { return fract(sin(n)*43758.5453123); }
float hash21(vec2 p) { p = fract(p*vec2(123.34, 456.21)); p += dot(p, p+45.32); return fract(p.x*p.y); }
float noise(in vec3 x) {
    vec3 p = floor(x); vec3 f = fract(x);
    f = f*f*(3.0-2.0*f);
    float n = p.x + p.y*57.0 + 113.0*p.z;
    return mix(mix(mix( hash(n+0.0), hash(n+1.0),f.x), mix( hash(n+57.0), hash(n+58.0),f.x),f.y),
               mix(mix( hash(n+113.0), hash(n+114.0),f.x), mix( hash(n+170.0), hash(n+171.0),f.x),f.y),f.z);
}
float sdSphere( vec3 p, float s ) { return length(p)-s; }
float sdBox( vec3 p, vec3 b ) { vec3 d = abs(p) - b; return min(max(d.x,max(d.y,d.z)),0.0) + length(max(d,0.0)); }
float sdTorus( vec3 p, vec2 t ) { return length( vec2(length(p.xz)-t.x,p.y) )-t.y; }
float sdCappedCone( in vec3 p, in float h, in float r1, in float r2 ) {
    vec2 q = vec2( length(p.xz), p.y );
    vec2 k1 = vec2(r2,h); vec2 k2 = vec2(r2-r1,2.0*h);
    vec2 ca = vec2(q.x-min(q.x,(q.y < 0.0)?r1:r2), abs(q.y)-h);
    vec2 cb = q - k1 + k2*clamp( dot(k1-q,k2)/dot2(k2), 0.0, 1.0 );
    float s = (cb.x < 0.0 && ca.y < 0.0) ? -1.0 : 1.0;
    return s*sqrt( min(dot2(ca),dot2(cb)) );
}
float smin( float a, float b, float k ) { float h = max(k-abs(a-b),0.0); return min(a, b) - h*h*0.25/k; }
vec2 opU( vec2 d1, vec2 d2 ) { return (d1.x<d2.x) ? d1 : d2; }
float opRep(inout vec3 p, float c) { float id = floor(p.x/c + .5); p.x = mod(p.x + .5*c, c) - .5*c; return id; }

Hit scene(vec3 p) {
    Hit h = Hit(FAR, -1);
    vec3 q = p;
    q.xz *= rot(0.3 + iTime);
    float id = opRep(q, 2.5);
    float d = sdBox(q, vec3(.5,.4,.3)) - .05;
    d = smin(d, sdSphere(q - vec3(0,.6,0), .35), .2);
    d = max(d, -sdTorus(q.xzy, vec2(.45,.1)));
    for (int i = ZERO; i < 3; ++i) {
        float fi = float(i);
        d = min(d, sdCappedCone(p - vec3(fi - 1., -.8, .5*fi), .3, .2, .05 + .02*hash(id)));
    }
    d += 0.02 * noise(p * 7.);
    if (d < h.d) { h.d = d; h.id = 1; }
    float g = p.y + 1.1;
    if (g < h.d) { h.d = g; h.id = 0; }
    return h;
}
float map(vec3 p) { return scene(p).d; }
vec3 calcNormal( in vec3 pos ) {
    vec3 n = vec3(0.0);
    for( int i=ZERO; i<4; i++ ) {
        vec3 e = 0.5773*(2.0*vec3((((i+3)>>1)&1),((i>>1)&1),(i&1))-1.0);
        n += e*map(pos+0.0005*e);
    }
    return normalize(n);
}
float softshadow(vec3 ro, vec3 rd, float mint, float tmax) {
    float res = 1.0, t = mint;
    for (int i = 0; i < 24; i++) { float h = map(ro + rd*t); res = min(res, 8.0*h/t); t += clamp(h, 0.02, 0.2); if (res < 0.004 || t > tmax) break; }
    return sat(res);
}
vec3 render(vec3 ro, vec3 rd) {
    float t = 0.; int id = -1;
    for (int i = 0; i < MAX_STEPS && t < FAR; i++) { Hit h = scene(ro + rd*t); if (abs(h.d) < EPS*t) { id = h.id; break; } t += h.d; }
    vec3 col = vec3(.7,.8,1.) - rd.y*.5;
    if (id >= 0) {
        vec3 p = ro + rd*t, n = calcNormal(p), l = normalize(vec3(.6,.7,-.5));
        float dif = sat(dot(n,l)) * softshadow(p, l, .02, 2.5);
        col = cols[id % 3] * (dif + .1);
        col = mix(col, vec3(.7,.8,1.), 1. - exp(-.0005*t*t*t));
    }
    return pow(col, vec3(.4545));
}
void mainImage( out vec4 fragColor, in vec2 fragCoord ) {
    vec2 uv = (2.*fragCoord - iResolution.xy)/iResolution.y;
    vec3 ro = vec3(0,1,-4), rd = normalize(vec3(uv, 1.5));
    fragColor = vec4(render(ro, rd), 1);
}
