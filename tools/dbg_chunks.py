import sys, numpy as np
sys.path.insert(0,'/root/repo')
import sdf2mesh_b200 as s2m, oracle
from tests.conftest import load_example_shader
ctx = s2m.Context(0)
mod = load_example_shader("mandelbulb").create_shader_module(ctx)
res=128
o = oracle.mesh_run("mandelbulb", res, 5.0)
for budget in [0, 129*160*4*20, 129*160*4*9, 129*160*4*3]:
    p,_ = s2m.params_from_cli(res, 5.0, flags=s2m.MESH_KEEP_CANDIDATES); p.slab_budget_bytes = budget
    for rep in range(2):
        r = s2m.mesh_run(ctx, mod, p); d = r.data()
        ok_keys = len(d.keys)==len(o.keys) and np.array_equal(d.keys,o.keys)
        ok_quads = len(d.quads)==len(o.quads) and np.array_equal(d.quads,o.quads)
        print("budget",budget,"rep",rep,"chunks",d.timings['chunks'],"cand",d.n_candidates,"nv",len(d.keys),len(o.keys),"nq",len(d.quads),len(o.quads),"inv",d.n_invalid_quads,o.n_invalid_quads,"keys",ok_keys,"quads",ok_quads)
        if not ok_keys:
            a=set(d.keys.tolist()); b=set(o.keys.tolist())
            extra=sorted(a-b)[:5]; missing=sorted(b-a)[:5]
            f=lambda k:(k&0xffff,(k>>16)&0xffff,k>>32)
            print("  extra",[f(k) for k in extra],"missing",[f(k) for k in missing])
        r.free()
