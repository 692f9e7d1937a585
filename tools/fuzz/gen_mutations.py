"""usage: gen_mutations.py SEED COUNT OUT_DIR -- damaged copies of the examples and tests/data fixtures (same mutations as
tests/test_frontend_fuzz.py::test_damaged_sources_give_errors_not_crashes)"""
import sys, os, random
ROOT=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0,ROOT)
from tests.test_frontend_fuzz import MUTATION_TOKENS
rnd=random.Random(int(sys.argv[1])); n=int(sys.argv[2]); out=sys.argv[3]
os.makedirs(out, exist_ok=True)
corpus=[(open(os.path.join(ROOT,"examples",f)).read(),"sdf3d") for f in ("torus.sdf3d","martin_cube.sdf3d","p_key.sdf3d")]
corpus.append((open(os.path.join(ROOT,"examples","mandelmesh.frag")).read(),"glsl"))
corpus.append((open(os.path.join(ROOT,"tests","data","wgsl_features.sdf3d")).read(),"sdf3d"))
for f in sorted(os.listdir(os.path.join(ROOT,"tests","data"))):
    if f.endswith(".glsl"):
        corpus.append(("#version 450 core\nuniform float iTime; uniform vec3 iResolution; uniform int iFrame; uniform vec4 iMouse;\n"+open(os.path.join(ROOT,"tests","data",f)).read()+"\nvoid main() {}\n","glsl"))
import json
corpus.append((json.dumps({"Shader": {"ver": "0.1", "info": {"id": "x", "name": "caf\u00e9 \U0001F600", "username": "u", "tags": ["a", "b"], "likes": 3},
                                      "renderpass": [{"inputs": [{"id": 1, "sampler": {"filter": "linear", "code": "no"}}], "outputs": [], "code": "float sdf(vec3 p) {\n\treturn length(p) - 0.75; // \"q\" \\ \n}\nvoid mainImage(out vec4 c, in vec2 u) { c = vec4(0.0); }\n", "name": "Image", "type": "image"}]}}), "json"))
corpus.append(('{"Error": "Shader not found"}', "json"))
for it in range(n):
    s,kind=rnd.choice(corpus)
    for _ in range(rnd.randint(0,4)):
        m,i=rnd.random(),rnd.randrange(len(s)+1)
        if m<0.3: s=s[:i]+s[min(len(s),i+rnd.randint(1,12)):]
        elif m<0.6: s=s[:i]+rnd.choice(MUTATION_TOKENS + ["\"", "\\u", "\\ud83d", "null", "true", "[", "]", "{\"code\":", ":"])+s[i:]
        elif m<0.75:
            j=min(len(s),i+rnd.randint(1,30)); s=s[:i]+s[i:j]*2+s[j:]
        elif m<0.9: s=s[:i]+" "+rnd.choice(MUTATION_TOKENS + ["\"", "\\u", "\\ud83d", "null", "true", "[", "]", "{\"code\":", ":"])+" "+s[i:]
        else: s=s[:i]
    open(os.path.join(out,"m%05d.%s"%(it,kind)),"w",encoding="utf-8",errors="surrogateescape").write(s)
